/* Single-rank MPI stand-in used ONLY to compile the unmodified BEEP reference
 * sources into oracle/_ref (test infrastructure, never part of the product).
 * No MPI is installed in the build image; BEEP uses 14 MPI symbols
 * (reference src/mpi.cc, utils.cc:323-335, mvn.cc:163, reader.cc:98), all of
 * which have a trivial meaning on one rank: gather/scatter are copies,
 * broadcast and barrier do nothing, and point-to-point is never reached
 * because a rank never sends to itself with ncore == 1. */
#ifndef BEEP_ORACLE_MPI_SHIM_H
#define BEEP_ORACLE_MPI_SHIM_H
#include <string.h>
#include <stdlib.h>
#include <stdio.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_DOUBLE 8
#define MPI_UNSIGNED 4
#define MPI_LONG 9
#define MPI_CXX_BOOL 1

static inline size_t beep_shim_sizeof(MPI_Datatype t)
{
	switch (t) {
	case MPI_DOUBLE: return sizeof(double);
	case MPI_UNSIGNED: return sizeof(unsigned int);
	case MPI_LONG: return sizeof(long);
	case MPI_CXX_BOOL: return sizeof(bool);
	}
	return 1;
}

static inline int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Gather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, int, MPI_Comm)
{
	if (r != s) memmove(r, s, (size_t)n * beep_shim_sizeof(t));
	return MPI_SUCCESS;
}
static inline int MPI_Scatter(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, int, MPI_Comm)
{
	if (r != s) memmove(r, s, (size_t)n * beep_shim_sizeof(t));
	return MPI_SUCCESS;
}
static inline int beep_shim_p2p(const char *what)
{
	fprintf(stderr, "mpi shim: %s reached with one rank\n", what);
	abort();
	return 1;
}
static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { return beep_shim_p2p("MPI_Send"); }
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { return beep_shim_p2p("MPI_Recv"); }
static inline int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return beep_shim_p2p("MPI_Isend"); }
static inline int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return beep_shim_p2p("MPI_Irecv"); }
static inline int MPI_Waitall(int, MPI_Request *, MPI_Status *) { return MPI_SUCCESS; }
#endif
