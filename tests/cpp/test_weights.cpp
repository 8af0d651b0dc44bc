// Host build of the product's per-texel function (3dworld_b200/csrc/tw_weights.cuh - the body of the CUDA kernel of tw_tile_weights_batch) so that its arithmetic can
// be checked against the oracle without a GPU (tests/test_weights_host.py). Not a CPU fallback: it is compiled by the test, never shipped in the library.
//   in:  uint32 ntiles, zvsize; tw_weight_params (class_ix filled); float zvals[nt*zv*zv]; float rand[nt*(zv-1)^2] (un-scaled noise); float tile_params[nt*8]
//   out: rgba[nt*(zv-1)^2*4], flags[nt]
#include "../../3dworld_b200/csrc/tw_weights.cuh"
#include <cstdio>
#include <cstdint>
#include <vector>

int main(int argc, char **argv) {
	if (argc < 3) return 1;
	FILE *in = fopen(argv[1], "rb"), *out = fopen(argv[2], "wb");
	if (!in || !out) return 1;
	uint32_t hdr[2];
	tw_weight_params W;
	if (fread(hdr, 4, 2, in) != 2 || fread(&W, sizeof(W), 1, in) != 1) return 1;
	uint32_t const nt = hdr[0], zv = hdr[1], st = zv - 1;
	std::vector<float> z((size_t)nt*zv*zv), r((size_t)nt*st*st), tp((size_t)nt*8);
	if (fread(z.data(), 4, z.size(), in) != z.size() || fread(r.data(), 4, r.size(), in) != r.size() || fread(tp.data(), 4, tp.size(), in) != tp.size()) return 1;
	std::vector<unsigned char> rgba((size_t)nt*st*st*4), flags(nt, 0);
	for (uint32_t t = 0; t < nt; ++t) {
		for (uint32_t y = 0; y < st; ++y) {
			for (uint32_t x = 0; x < st; ++x) {
				float const rand_offset = W.noise_scale*r[((size_t)t*st + y)*st + x];
				if (tww::weights_texel(z.data() + (size_t)t*zv*zv, zv, x, y, rand_offset, tp.data() + (size_t)t*8, W, &rgba[(((size_t)t*st + y)*st + x)*4])) {flags[t] = 1;}
			}
		}
	}
	fwrite(rgba.data(), 1, rgba.size(), out); fwrite(flags.data(), 1, flags.size(), out);
	fclose(in); fclose(out);
	return 0;
}
