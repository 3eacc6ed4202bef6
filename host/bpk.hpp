// bpk.hpp -- reader of the "BPK1" named-array container (the model description the reference-side
// adapter writes, oracle/ref/ref_harness.cc:export_desc; layout in beep_b200/bpk.py) into a BeepModelDesc.
#ifndef BEEP_BPK_HPP
#define BEEP_BPK_HPP
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "beep_b200.h"

namespace beep {

struct BpkArray { uint32_t dtype; std::vector<int64_t> dims; std::vector<unsigned char> data; size_t count() const { size_t n = 1; for (auto d : dims) n *= (size_t)d; return n; } };

inline std::map<std::string, BpkArray> read_bpk(const std::string &path)
{
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) throw std::runtime_error("cannot open " + path);
	char magic[4];
	uint32_t n = 0;
	if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "BPK1", 4) != 0 || fread(&n, 4, 1, f) != 1) { fclose(f); throw std::runtime_error(path + ": not a BPK1 file"); }
	std::map<std::string, BpkArray> out;
	const size_t esz[4] = {4, 8, 8, 1};
	for (uint32_t i = 0; i < n; i++) {
		uint32_t len = 0, nd = 0;
		BpkArray a;
		if (fread(&len, 4, 1, f) != 1) break;
		std::string name(len, ' ');
		if (fread(&name[0], 1, len, f) != len || fread(&a.dtype, 4, 1, f) != 1 || fread(&nd, 4, 1, f) != 1 || a.dtype > 3) { fclose(f); throw std::runtime_error(path + ": corrupt"); }
		a.dims.resize(nd);
		if (nd && fread(a.dims.data(), 8, nd, f) != nd) { fclose(f); throw std::runtime_error(path + ": corrupt"); }
		a.data.resize(a.count() * esz[a.dtype]);
		if (!a.data.empty() && fread(a.data.data(), 1, a.data.size(), f) != a.data.size()) { fclose(f); throw std::runtime_error(path + ": truncated"); }
		out[name] = std::move(a);
	}
	fclose(f);
	return out;
}

/// Owns the arrays a BeepModelDesc points into.
struct ModelFile {
	std::map<std::string, BpkArray> arrays;
	BeepModelDesc desc;
	explicit ModelFile(const std::string &path, const std::string &prefix = "desc.") : arrays(read_bpk(path))
	{
		memset(&desc, 0, sizeof desc);
		auto get = [&](const char *name, uint32_t dtype) -> BpkArray & {
			auto it = arrays.find(prefix + name);
			if (it == arrays.end() || it->second.dtype != dtype) throw std::runtime_error(std::string("model file lacks field ") + name);
			return it->second;
		};
#define BEEP_I32(name) desc.name = *(const int32_t *)get(#name, 0).data.data();
#define BEEP_F64(name) desc.name = *(const double *)get(#name, 1).data.data();
#define BEEP_AI32(name) { BpkArray &a = get(#name, 0); desc.name = (const int32_t *)a.data.data(); desc.n_##name = (int64_t)a.count(); }
#define BEEP_AF64(name) { BpkArray &a = get(#name, 1); desc.name = (const double *)a.data.data(); desc.n_##name = (int64_t)a.count(); }
#include "beep_model_fields.def"
#undef BEEP_I32
#undef BEEP_F64
#undef BEEP_AI32
#undef BEEP_AF64
	}
};

} // namespace beep
#endif
