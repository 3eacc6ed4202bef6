/* comm.cuh -- the collectives of the particle methods inside the C-ABI library, NCCL over NVLink, device to device.
 *
 * Replaces the particle plumbing of the reference's Mpi class (one MPI rank per core there, one process -- or one
 * BeepComm -- per GPU here):
 *     Mpi::gather / bcast of per-particle scalars     src/mpi.cc:393-676   -> beep_allgather_scalar (ncclAllGather)
 *     Mpi::copy_particles                              src/mpi.cc:81-165    -> beep_copy_particles
 *     Mpi::end_sec_swap                                src/mpi.cc:215-327   -> beep_end_sec_swap
 *     Mpi::particle_sample                             src/mpi.cc:331-389   -> beep_particle_sample
 * Global particle i lives on rank i / N (N = particles per rank), the reference's core = i/N (mpi.cc:98-101).
 * The decisions (who copies whom) are replicated on every rank (beep_bootstrap, beep_ef_cutoff, beep_pas_bootstrap,
 * beep_mc3_swap from a shared key), so every rank derives the same send / receive lists from the same vector and no
 * broadcast is needed.  Particles move as the contiguous per-particle slabs of the batch (parameters, derived tables,
 * event counts, means, population, observation sums) straight from HBM to HBM -- grouped ncclSend / ncclRecv; what
 * arrives is staged and copied into place after every source has been read, the reference's send-then-overwrite order.
 *
 * Included at the end of engine.cu (one translation unit: the structs of the engine are private to it).
 */
#ifndef BEEP_COMM_CUH
#define BEEP_COMM_CUH

#include <nccl.h>
#include <map>

struct BeepComm {
	BeepEngine *eng = nullptr;
	ncclComm_t comm = nullptr;
	int rank = 0, nranks = 1;
	cudaStream_t stream = nullptr;
	char *stage = nullptr; size_t stage_cap = 0;      /* grow-only staging of what arrives */
	double *d_gather = nullptr; size_t gather_cap = 0;
	long long bytes_sent = 0, bytes_recv = 0;         /* over the communicator's life (bench accounting) */
};

#define NC(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) return fail(BEEP_ERR_CUDA, "%s: %s (%s:%d)", #x, ncclGetErrorString(r_), __FILE__, __LINE__); } while (0)

extern "C" int beep_comm_unique_id(void *id128)
{
	if (!id128) return fail(BEEP_ERR_INVALID, "null argument");
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	NC(ncclGetUniqueId(&id));
	memcpy(id128, &id, sizeof id);
	return BEEP_OK;
}

extern "C" void beep_comm_destroy(BeepComm *cm)
{
	if (!cm) return;
	if (cm->eng) cudaSetDevice(cm->eng->device);
	if (cm->stream) cudaStreamSynchronize(cm->stream);
	if (cm->comm) ncclCommDestroy(cm->comm);
	if (cm->stage) cudaFree(cm->stage);
	if (cm->d_gather) cudaFree(cm->d_gather);
	if (cm->stream) cudaStreamDestroy(cm->stream);
	delete cm;
}

extern "C" int beep_comm_create(BeepEngine *eng, const void *id128, int nranks, int rank, BeepComm **out)
{
	if (!eng || !out || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) return fail(BEEP_ERR_INVALID, "bad argument to beep_comm_create");
	*out = nullptr;
	CU(cudaSetDevice(eng->device));
	BeepComm *cm = new BeepComm();
	cm->eng = eng; cm->rank = rank; cm->nranks = nranks;
	cudaError_t e = cudaStreamCreateWithFlags(&cm->stream, cudaStreamNonBlocking);
	if (e != cudaSuccess) { delete cm; return fail(BEEP_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
	if (nranks > 1) {
		ncclUniqueId id;
		memcpy(&id, id128, sizeof id);
		ncclResult_t r = ncclCommInitRank(&cm->comm, nranks, id, rank);
		if (r != ncclSuccess) { beep_comm_destroy(cm); return fail(BEEP_ERR_CUDA, "ncclCommInitRank: %s", ncclGetErrorString(r)); }
	}
	*out = cm;
	return BEEP_OK;
}
extern "C" int beep_comm_rank(const BeepComm *cm) { return cm ? cm->rank : 0; }
extern "C" int beep_comm_size(const BeepComm *cm) { return cm ? cm->nranks : 1; }
extern "C" int64_t beep_comm_bytes_moved(const BeepComm *cm, int64_t *recv) { if (recv) *recv = cm ? cm->bytes_recv : 0; return cm ? cm->bytes_sent : 0; }

/* per-particle scalars of every rank, in rank order (8 n_local bytes per rank): device arrays, on `stream` */
extern "C" int beep_allgather_dev(BeepComm *cm, const double *local_dev, int n_local, double *all_dev, void *stream)
{
	if (!cm || !local_dev || !all_dev || n_local <= 0) return fail(BEEP_ERR_INVALID, "bad argument to beep_allgather_dev");
	CU(cudaSetDevice(cm->eng->device));
	cudaStream_t st = stream ? (cudaStream_t)stream : cm->stream;
	if (cm->nranks == 1) { CU(cudaMemcpyAsync(all_dev, local_dev, (size_t)n_local * sizeof(double), cudaMemcpyDeviceToDevice, st)); return BEEP_OK; }
	NC(ncclAllGather(local_dev, all_dev, (size_t)n_local, ncclDouble, cm->comm, st));
	return BEEP_OK;
}

/* replaces mpi.gather + the decision on rank 0 + mpi.bcast (abcmbp.cc:141, pas.cc:133, pmcmc.cc:219,263,
 * mc3.cc:170-175) by one all-gather; host arrays in and out */
extern "C" int beep_allgather_scalar(BeepComm *cm, const double *local_host, int n_local, double *all_host)
{
	if (!cm || !local_host || !all_host || n_local <= 0) return fail(BEEP_ERR_INVALID, "bad argument to beep_allgather_scalar");
	if (cm->nranks == 1) { memcpy(all_host, local_host, (size_t)n_local * sizeof(double)); return BEEP_OK; }
	CU(cudaSetDevice(cm->eng->device));
	const size_t need = (size_t)n_local * (cm->nranks + 1);
	if (cm->gather_cap < need) {
		if (cm->d_gather) cudaFree(cm->d_gather);
		cm->d_gather = nullptr; cm->gather_cap = 0;
		CU(cudaMalloc((void **)&cm->d_gather, need * sizeof(double)));
		cm->gather_cap = need;
	}
	double *d_loc = cm->d_gather, *d_all = cm->d_gather + n_local;
	CU(cudaMemcpyAsync(d_loc, local_host, (size_t)n_local * sizeof(double), cudaMemcpyHostToDevice, cm->stream));
	NC(ncclAllGather(d_loc, d_all, (size_t)n_local, ncclDouble, cm->comm, cm->stream));
	CU(cudaMemcpyAsync(all_host, d_all, (size_t)n_local * cm->nranks * sizeof(double), cudaMemcpyDeviceToHost, cm->stream));
	CU(cudaStreamSynchronize(cm->stream));
	return BEEP_OK;
}

/* ------------------------------------------------------------------------ */
/* the per-particle slabs of a batch                                          */
/* ------------------------------------------------------------------------ */
struct Slab { char *base; size_t bytes; };
static void batch_slabs(BeepBatch *bt, bool full, std::vector<Slab> &v)
{
	const DevModel &m = bt->eng->dm;
	const DevBatch &b = bt->db;
	auto add = [&](void *p, size_t n, size_t el) { if (p && n) v.push_back(Slab{(char *)p, n * el}); };
	add(b.pop_cur, (size_t)m.A * m.CDp, 8); add(b.popi, (size_t)m.A * m.CDp4, 4);
	if (!full) return;
	const size_t TAJD = (size_t)m.T * m.A * m.J * m.D;
	add(b.theta, m.nparam, 8); add(b.pd, m.nparam, 8); add(b.transrate, (size_t)m.J * m.D, 8); add(b.disc_spline, (size_t)m.nspline * m.T, 8);
	add(b.sus, m.D, 8); add(b.areafactor, (size_t)m.period * m.A, 8); add(b.betaR, (size_t)m.S * m.nR * m.T, 8); add(b.infco, m.C, 8);
	add(b.infmask, 1, 8); add(b.etab, (size_t)m.nefoi * m.T, 8); add(b.crate, ((size_t)m.J * m.D + 31) & ~(size_t)31, 8); add(b.inflist, (size_t)m.C + 1, 4);
	add(b.tn, TAJD, 4); add(b.tm, TAJD, 8); add(b.poptraj, (size_t)m.T * m.A * m.C * m.D, 8);
	add(b.obs_partial, m.nslot, 8); add(b.obs_val, m.nobs, 8);
}

struct MovePlan {
	std::vector<uint32_t> local;                        /* [N] local source slot or 0xFFFFFFFF */
	struct Out { int peer; int slot; };                 /* send the content of local `slot` to rank `peer` */
	struct In { int peer; std::vector<int> slots; };    /* receive one particle from `peer`, copy it to these local slots */
	std::vector<Out> out;
	std::vector<In> in;
};

/* Mpi::copy_particles (mpi.cc:81-165): global particle i continues as a copy of partcopy[i] (0xFFFFFFFF = keep).
 * Sends are listed by ascending global destination, receives by ascending local slot (= global destination), so the
 * k-th message between a pair of ranks is the same particle on both sides. */
static void plan_copy(const uint32_t *partcopy, int ntot, int N, int rank, MovePlan &pl)
{
	const int lo = rank * N;
	pl.local.assign(N, 0xFFFFFFFFu);
	for (int i = 0; i < ntot; i++) {
		const uint32_t s = partcopy[i];
		if (s == 0xFFFFFFFFu || (int)s == i) continue;
		const bool src_here = (int)s >= lo && (int)s < lo + N, dst_here = i >= lo && i < lo + N;
		if (src_here && dst_here) pl.local[i - lo] = s - lo;
		else if (src_here) pl.out.push_back(MovePlan::Out{i / N, (int)s - lo});
		else if (dst_here) pl.in.push_back(MovePlan::In{(int)s / N, std::vector<int>(1, i - lo)});
	}
}

/* Mpi::end_sec_swap (mpi.cc:215-327): particle p continues from the state of backpart[p].  Only DISTINCT remote
 * ancestors travel (preclist, mpi.cc:243-258; one message per (ancestor, destination rank), corlist :270-287);
 * both sides order the messages of a pair of ranks by ascending ancestor. */
static void plan_section_swap(const uint32_t *backpart, int ntot, int N, int rank, MovePlan &pl)
{
	const int lo = rank * N;
	pl.local.assign(N, 0xFFFFFFFFu);
	std::map<int, std::vector<int>> need;               /* remote ancestor -> local slots, ascending ancestor */
	for (int k = 0; k < N; k++) {
		const int pp = (int)backpart[lo + k];
		if (backpart[lo + k] == 0xFFFFFFFFu || pp == lo + k) continue;
		if (pp >= lo && pp < lo + N) pl.local[k] = (uint32_t)(pp - lo);
		else need[pp].push_back(k);
	}
	for (auto &kv : need) pl.in.push_back(MovePlan::In{kv.first / N, kv.second});
	/* one pass over the ancestry (the reference scans all Ntot particles for each of its own, mpi.cc:264-287):
	 * the ranks that need each local particle */
	const int nranks = ntot / N;
	std::vector<std::vector<unsigned char>> want(N);
	for (int pp = 0; pp < ntot; pp++) {
		const int p = (int)backpart[pp], cor = pp / N;
		if (backpart[pp] == 0xFFFFFFFFu || p < lo || p >= lo + N || cor == rank) continue;
		if (want[p - lo].empty()) want[p - lo].assign(nranks, 0);
		want[p - lo][cor] = 1;
	}
	for (int k = 0; k < N; k++)
		if (!want[k].empty())
			for (int cor = 0; cor < nranks; cor++) if (want[k][cor]) pl.out.push_back(MovePlan::Out{cor, k});
}

/* test hook: the plan of one rank as flat arrays (tests/test_comm_plan_cpu.py replays the exchange of every rank)
 *   kind 0 = copy_particles, 1 = end_sec_swap.  local[N]; out as (peer, slot) pairs; in as (peer, nslots, slots...) */
extern "C" int beep_test_move_plan(int kind, const uint32_t *vec_all, int ntot, int N, int rank, uint32_t *local, int32_t *out_pairs, int *nout,
                                   int32_t *in_flat, int *nin_words)
{
	if (!vec_all || ntot <= 0 || N <= 0 || ntot % N || rank < 0 || rank >= ntot / N || !local || !out_pairs || !nout || !in_flat || !nin_words) return fail(BEEP_ERR_INVALID, "bad argument to beep_test_move_plan");
	MovePlan pl;
	if (kind == 0) plan_copy(vec_all, ntot, N, rank, pl); else plan_section_swap(vec_all, ntot, N, rank, pl);
	for (int k = 0; k < N; k++) local[k] = pl.local[k];
	*nout = (int)pl.out.size();
	for (size_t k = 0; k < pl.out.size(); k++) { out_pairs[2 * k] = pl.out[k].peer; out_pairs[2 * k + 1] = pl.out[k].slot; }
	int w = 0;
	for (auto &in : pl.in) { in_flat[w++] = in.peer; in_flat[w++] = (int)in.slots.size(); for (int s : in.slots) in_flat[w++] = s; }
	*nin_words = w;
	return BEEP_OK;
}

/* pack / unpack of small per-particle slabs: row k of `dst` <- row idx[k] of `src` (pack), row dst_idx[k] of `dst`
 * <- row src_idx[k] of `src` (unpack).  n = 4-byte words per row. */
__global__ void k_rows_gather(unsigned *dst, const unsigned *src, const int *idx, size_t n)
{
	const unsigned *sp = src + (size_t)idx[blockIdx.x] * n;
	unsigned *dp = dst + (size_t)blockIdx.x * n;
	for (size_t i = blockIdx.y * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.y * blockDim.x) dp[i] = sp[i];
}
__global__ void k_rows_scatter(unsigned *dst, const unsigned *src, const int *src_idx, const int *dst_idx, size_t n)
{
	const unsigned *sp = src + (size_t)src_idx[blockIdx.x] * n;
	unsigned *dp = dst + (size_t)dst_idx[blockIdx.x] * n;
	for (size_t i = blockIdx.y * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.y * blockDim.x) dp[i] = sp[i];
}

/* carry out a plan: sends from the particles' slabs, receives into the staging area; then the copies within the rank
 * (out of place: every source is read before any destination is written); then the staged particles go to their slots.
 *   full  (whole particles, MB-sized slabs): grouped ncclSend / ncclRecv straight from / next to the slabs;
 *   !full (the section state of the particle filter, ~1 KB per particle, hundreds of particles per section): the
 *         outgoing rows are PACKED per peer by one kernel, ONE ncclSend and ONE ncclRecv per peer and slab move them,
 *         one kernel scatters what arrived -- a handful of launches per section instead of one NCCL call and one copy
 *         per particle (measured: 5.4 ms -> see profiles/README.md per section at 2 x 2 048 particles). */
static int run_plan(BeepComm *cm, BeepBatch *bt, const MovePlan &pl, bool full)
{
	std::vector<Slab> slabs;
	batch_slabs(bt, full, slabs);
	size_t per = 0;
	for (auto &s : slabs) per += (s.bytes + 255) & ~(size_t)255;
	cudaStream_t st = bt->stream;
	if ((!pl.in.empty() || !pl.out.empty()) && cm->nranks < 2) return fail(BEEP_ERR_INVALID, "the plan moves particles between ranks but the communicator has one");
	bool any_local = false;
	for (uint32_t v : pl.local) if (v != 0xFFFFFFFFu) any_local = true;
	const bool packed = !full;
	for (auto &sl : slabs) if (sl.bytes % 4) return fail(BEEP_ERR_INVALID, "slab of %zu bytes", sl.bytes);
	if (packed) {
		/* messages of a pair of ranks keep the plan's order (ascending ancestor on both sides) */
		const int nr = cm->nranks;
		std::vector<std::vector<int>> out_of(nr), in_of(nr);          /* per peer: local slots to send / indices into pl.in */
		for (auto &o : pl.out) out_of[o.peer].push_back(o.slot);
		for (size_t j = 0; j < pl.in.size(); j++) in_of[pl.in[j].peer].push_back((int)j);
		std::vector<int> idx;                                           /* [out rows | unpack source rows | unpack destination slots] */
		std::vector<size_t> out_off(nr + 1, 0), in_off(nr + 1, 0);
		for (int r = 0; r < nr; r++) { out_off[r + 1] = out_off[r] + out_of[r].size(); in_off[r + 1] = in_off[r] + in_of[r].size(); }
		const size_t nout = out_off[nr], nin = in_off[nr];
		for (int r = 0; r < nr; r++) for (int sl : out_of[r]) idx.push_back(sl);
		std::vector<int> usrc, udst;
		for (int r = 0; r < nr; r++)
			for (size_t k = 0; k < in_of[r].size(); k++)
				for (int slot : pl.in[in_of[r][k]].slots) { usrc.push_back((int)(in_off[r] + k)); udst.push_back(slot); }
		const size_t npair = usrc.size();
		idx.insert(idx.end(), usrc.begin(), usrc.end());
		idx.insert(idx.end(), udst.begin(), udst.end());
		const size_t idx_bytes = (idx.size() * sizeof(int) + 255) & ~(size_t)255;
		const size_t need = per * (nout + nin) + idx_bytes;
		if (need > cm->stage_cap) {
			CU(cudaStreamSynchronize(st));
			CU(cudaStreamSynchronize(cm->stream));
			if (cm->stage) cudaFree(cm->stage);
			cm->stage = nullptr; cm->stage_cap = 0;
			CU(cudaMalloc((void **)&cm->stage, need + need / 2));
			cm->stage_cap = need + need / 2;
		}
		int *d_idx = (int *)cm->stage;
		char *buf = cm->stage + idx_bytes;
		if (!idx.empty()) CU(cudaMemcpyAsync(d_idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice, st));
		/* per slab: [send rows of all peers | receive rows of all peers] */
		std::vector<char *> sbuf(slabs.size()), rbuf(slabs.size());
		{
			char *q = buf;
			for (size_t k = 0; k < slabs.size(); k++) { const size_t rb = (slabs[k].bytes + 255) & ~(size_t)255; sbuf[k] = q; q += rb * nout; rbuf[k] = q; q += rb * nin; }
		}
		if (nout)
			for (size_t k = 0; k < slabs.size(); k++) {
				const size_t w = slabs[k].bytes / 4;
				k_rows_gather<<<dim3((unsigned)nout, (unsigned)std::max<size_t>(1, std::min<size_t>(8, w / 256))), 256, 0, st>>>((unsigned *)sbuf[k], (const unsigned *)slabs[k].base, d_idx, w);
			}
		if (nout || nin) {
			NC(ncclGroupStart());
			for (int r = 0; r < nr; r++)
				for (size_t k = 0; k < slabs.size(); k++) {
					const size_t b = slabs[k].bytes;
					if (!out_of[r].empty()) { NC(ncclSend(sbuf[k] + out_off[r] * b, out_of[r].size() * b, ncclChar, r, cm->comm, st)); cm->bytes_sent += (long long)(out_of[r].size() * b); }
					if (!in_of[r].empty()) { NC(ncclRecv(rbuf[k] + in_off[r] * b, in_of[r].size() * b, ncclChar, r, cm->comm, st)); cm->bytes_recv += (long long)(in_of[r].size() * b); }
				}
			NC(ncclGroupEnd());
		}
		if (any_local) {
			int rc = beep_permute(bt, pl.local.data(), BEEP_PERMUTE_SECTION_STATE);
			if (rc) return rc;
		}
		if (npair)
			for (size_t k = 0; k < slabs.size(); k++) {
				const size_t w = slabs[k].bytes / 4;
				k_rows_scatter<<<dim3((unsigned)npair, (unsigned)std::max<size_t>(1, std::min<size_t>(8, w / 256))), 256, 0, st>>>((unsigned *)slabs[k].base, (const unsigned *)rbuf[k], d_idx + nout, d_idx + nout + npair, w);
			}
		bt->eng->launches += (nout ? (long long)slabs.size() : 0) + (npair ? (long long)slabs.size() : 0);
		CU(cudaGetLastError());
		CU(cudaStreamSynchronize(st));
		return BEEP_OK;
	}
	const size_t need = per * pl.in.size();
	if (need > cm->stage_cap) {
		CU(cudaStreamSynchronize(cm->stream));
		if (cm->stage) cudaFree(cm->stage);
		cm->stage = nullptr; cm->stage_cap = 0;
		CU(cudaMalloc((void **)&cm->stage, need));
		cm->stage_cap = need;
	}
	if (!pl.in.empty() || !pl.out.empty()) {
		NC(ncclGroupStart());
		for (auto &o : pl.out)
			for (auto &s : slabs) { NC(ncclSend(s.base + (size_t)o.slot * s.bytes, s.bytes, ncclChar, o.peer, cm->comm, st)); cm->bytes_sent += (long long)s.bytes; }
		size_t off = 0;
		for (auto &in : pl.in)
			for (auto &s : slabs) { NC(ncclRecv(cm->stage + off, s.bytes, ncclChar, in.peer, cm->comm, st)); off += (s.bytes + 255) & ~(size_t)255; cm->bytes_recv += (long long)s.bytes; }
		NC(ncclGroupEnd());
	}
	if (any_local) {
		int rc = beep_permute(bt, pl.local.data(), BEEP_PERMUTE_FULL);
		if (rc) return rc;
	}
	size_t off = 0;
	for (auto &in : pl.in)
		for (auto &s : slabs) {
			for (int slot : in.slots) CU(cudaMemcpyAsync(s.base + (size_t)slot * s.bytes, cm->stage + off, s.bytes, cudaMemcpyDeviceToDevice, st));
			off += (s.bytes + 255) & ~(size_t)255;
		}
	CU(cudaStreamSynchronize(st));
	return BEEP_OK;
}

static int comm_batch_ok(BeepComm *cm, BeepBatch *bt, int ntot)
{
	if (!cm || !bt) return fail(BEEP_ERR_INVALID, "null argument");
	if (cm->eng != bt->eng) return fail(BEEP_ERR_INVALID, "communicator and batch belong to different engines");
	if (ntot != bt->db.P * cm->nranks) return fail(BEEP_ERR_INVALID, "%d particles in all, but %d ranks x %d per rank", ntot, cm->nranks, bt->db.P);
	return BEEP_OK;
}

/* Replaces Mpi::copy_particles (mpi.cc:81-165) across the ranks of the communicator: whole particles, exact state
 * (no replay needed afterwards). */
extern "C" int beep_copy_particles(BeepComm *cm, BeepBatch *bt, const uint32_t *partcopy_all_host, int ntot)
{
	int rc = comm_batch_ok(cm, bt, ntot);
	if (rc) return rc;
	if (!partcopy_all_host) return fail(BEEP_ERR_INVALID, "null argument");
	for (int i = 0; i < ntot; i++) if (partcopy_all_host[i] != 0xFFFFFFFFu && partcopy_all_host[i] >= (unsigned)ntot) return fail(BEEP_ERR_INVALID, "partcopy[%d] = %u out of range", i, partcopy_all_host[i]);
	CU(cudaSetDevice(bt->eng->device));
	if (bt->theta_dirty && bt->have_params && (rc = launch_set_param(bt))) return rc;
	MovePlan pl;
	plan_copy(partcopy_all_host, ntot, bt->db.P, cm->rank, pl);
	rc = run_plan(cm, bt, pl, true);
	if (rc) return rc;
	if (!pl.in.empty()) all_slots_have_theta(bt);
	return BEEP_OK;
}

/* Replaces Mpi::end_sec_swap (mpi.cc:215-327): the section state (population; the infectivity map is re-derived
 * from it) of every particle's ancestor. */
extern "C" int beep_end_sec_swap(BeepComm *cm, BeepBatch *bt, const uint32_t *backpart_all_host, int ntot)
{
	int rc = comm_batch_ok(cm, bt, ntot);
	if (rc) return rc;
	if (!backpart_all_host) return fail(BEEP_ERR_INVALID, "null argument");
	for (int i = 0; i < ntot; i++) if (backpart_all_host[i] != 0xFFFFFFFFu && backpart_all_host[i] >= (unsigned)ntot) return fail(BEEP_ERR_INVALID, "backpart[%d] = %u out of range", i, backpart_all_host[i]);
	CU(cudaSetDevice(bt->eng->device));
	MovePlan pl;
	plan_section_swap(backpart_all_host, ntot, bt->db.P, cm->rank, pl);
	return run_plan(cm, bt, pl, false);
}

/* Replaces Mpi::particle_sample (mpi.cc:331-389): walks the ancestry of the filter backwards from particle
 * backpart[nsec-1][0] and assembles ONE trajectory from the sections of the particles it passes through
 * (backpart_host is [nsec][Ntot]: row s = the ancestors chosen by the bootstrap after section s, i.e. the
 * reference's backpart[s+1]).  Rank 0 receives the pieces (device to device) and returns the trajectory as
 * doubles [T][A][J][D] with the batch's parameters of slot 0 (the filter's particles share them); the other
 * ranks send what they own and leave the outputs untouched.  cm may be NULL on a single GPU. */
extern "C" int beep_particle_sample(BeepComm *cm, BeepBatch *bt, const uint32_t *backpart_host, int nsec, int ntot, double *theta_host, double *transnum_host)
{
	if (!bt || !backpart_host) return fail(BEEP_ERR_INVALID, "null argument");
	BeepEngine *eng = bt->eng;
	const DevModel &m = eng->dm;
	const int N = bt->db.P, rank = cm ? cm->rank : 0, nranks = cm ? cm->nranks : 1;
	if (nsec != m.nsection) return fail(BEEP_ERR_INVALID, "the model has %d sections, not %d", m.nsection, nsec);
	if (ntot != N * nranks) return fail(BEEP_ERR_INVALID, "%d particles in all, but %d ranks x %d per rank", ntot, nranks, N);
	if (!m.stochastic || !bt->db.tn) return fail(BEEP_ERR_STATE, "beep_particle_sample needs a stochastic batch that keeps transnum");
	if (rank == 0 && !transnum_host) return fail(BEEP_ERR_INVALID, "rank 0 needs the output array");
	for (size_t i = 0; i < (size_t)nsec * ntot; i++) if (backpart_host[i] >= (unsigned)ntot) return fail(BEEP_ERR_INVALID, "backpart[%zu] = %u out of range", i, backpart_host[i]);
	CU(cudaSetDevice(eng->device));
	const size_t AJD = (size_t)m.A * m.J * m.D, TAJD = (size_t)m.T * AJD;
	int *d_traj = nullptr;
	if (rank == 0) CU(cudaMalloc((void **)&d_traj, TAJD * sizeof(int)));
	cudaStream_t st = bt->stream;
	bool grouped = false;
	unsigned p = backpart_host[(size_t)(nsec - 1) * ntot + 0];
	for (int sec = nsec - 1; sec >= 0; sec--) {
		const int ti = eng->section_ti[sec], tf = eng->section_tf[sec];
		const size_t n = (size_t)(tf - ti) * AJD;
		const int co = (int)p / N;
		const int *src = bt->db.tn + (size_t)(p % N) * TAJD + (size_t)ti * AJD;
		if (rank == 0) {
			if (co == 0) CU(cudaMemcpyAsync(d_traj + (size_t)ti * AJD, src, n * sizeof(int), cudaMemcpyDeviceToDevice, st));
			else { if (!grouped) { NC(ncclGroupStart()); grouped = true; } NC(ncclRecv(d_traj + (size_t)ti * AJD, n, ncclInt, co, cm->comm, st)); cm->bytes_recv += (long long)(n * sizeof(int)); }
		} else if (co == rank) {
			if (!grouped) { NC(ncclGroupStart()); grouped = true; }
			NC(ncclSend(src, n, ncclInt, 0, cm->comm, st));
			cm->bytes_sent += (long long)(n * sizeof(int));
		}
		if (sec > 0) p = backpart_host[(size_t)(sec - 1) * ntot + p];
	}
	if (grouped) NC(ncclGroupEnd());
	int rc = BEEP_OK;
	if (rank == 0) {
		std::vector<int> tmp(TAJD);
		cudaError_t e = cudaMemcpyAsync(tmp.data(), d_traj, TAJD * sizeof(int), cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaStreamSynchronize(st);
		cudaFree(d_traj);
		if (e != cudaSuccess) return fail(BEEP_ERR_CUDA, "beep_particle_sample: %s", cudaGetErrorString(e));
		for (size_t i = 0; i < TAJD; i++) transnum_host[i] = (double)tmp[i];
		if (theta_host) rc = beep_export_particle(bt, 0, theta_host, nullptr);
	} else CU(cudaStreamSynchronize(st));
	return rc;
}

#endif
