// tw_streaming.cu - HBM-bound element-wise passes on the far side of the height path:
//   heightmap_t::from_floats / to_floats 16-bit pack (src/heightmap.cpp:191-215, texture_t::write_pixel_16_bits src/Textures.cpp:1889-1893).
// One float4 (16 B) in, 8 B out per thread; grid sized to a multiple of the SM count.
#include "tw_internal.h"

namespace {

__device__ __forceinline__ unsigned pack16(float h, float val_add, float val_div, unsigned &bad) {
	float const v = (h - val_add)*val_div;                 // src/heightmap.cpp:210
	if (!(v >= 0.0f && v < 256.0f)) {bad = 1; return 0;}   // the reference asserts here (:211)
	unsigned const high_bits = (unsigned)v & 0xffu;        // (unsigned char)val - truncate
	unsigned const low_bits  = (unsigned)(256.0f*(v - (float)high_bits)) & 0xffu;
	return low_bits | (high_bits << 8);                    // data[2i] = low, data[2i+1] = high
}

__global__ void from_floats_u16_kernel(const float *__restrict__ vals, size_t n, float val_add, float val_div, uint16_t *__restrict__ out, unsigned *__restrict__ bad_count) {
	size_t const stride = (size_t)gridDim.x*blockDim.x;
	unsigned bad = 0;
	size_t const n4 = n/4;
	bool const aligned = ((((uintptr_t)vals) & 15) == 0) && ((((uintptr_t)out) & 7) == 0);
	if (aligned) {
		for (size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x; i < n4; i += stride) {
			float4 const v = __ldg(reinterpret_cast<const float4 *>(vals) + i);
			unsigned const a = pack16(v.x, val_add, val_div, bad) | (pack16(v.y, val_add, val_div, bad) << 16);
			unsigned const b = pack16(v.z, val_add, val_div, bad) | (pack16(v.w, val_add, val_div, bad) << 16);
			reinterpret_cast<uint2 *>(out)[i] = make_uint2(a, b);
		}
		for (size_t i = 4*n4 + (size_t)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += stride) {out[i] = (uint16_t)pack16(__ldg(vals + i), val_add, val_div, bad);}
	}
	else {
		for (size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += stride) {out[i] = (uint16_t)pack16(__ldg(vals + i), val_add, val_div, bad);}
	}
	if (bad) {atomicAdd(bad_count, 1u);}
}

__global__ void to_floats_u16_kernel(const uint16_t *__restrict__ data, size_t n, float val_mult, float val_add, float *__restrict__ vals) {
	size_t const stride = (size_t)gridDim.x*blockDim.x;
	for (size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += stride) {
		unsigned const p = __ldg(data + i);
		float const v = (float)((double)(p & 0xffu)/256.0 + (double)(p >> 8)); // data[i<<1]/256.0 + data[(i<<1)+1], src/heightmap.cpp:199
		vals[i] = val_mult*v + val_add;
	}
}

int grid_for(size_t n) {size_t b = (n + 1023)/1024; if (b > 148*16) b = 148*16; if (b < 1) b = 1; return (int)b;}

} // namespace

int twi_from_floats_u16(tw_ctx *ctx, const float *d_vals, size_t n, float val_mult, float val_add, uint8_t *d_out, unsigned *d_bad) {
	float const val_div = (float)(1.0/(double)val_mult); // src/heightmap.cpp:206
	from_floats_u16_kernel<<<grid_for(n), 256, 0, ctx->stream>>>(d_vals, n, val_add, val_div, reinterpret_cast<uint16_t *>(d_out), d_bad);
	TW_LAUNCH_CHECK(ctx);
	return TW_OK;
}

int twi_to_floats_u16(tw_ctx *ctx, const uint8_t *d_data, size_t n, float val_mult, float val_add, float *d_vals) {
	to_floats_u16_kernel<<<grid_for(n), 256, 0, ctx->stream>>>(reinterpret_cast<const uint16_t *>(d_data), n, val_mult, val_add, d_vals);
	TW_LAUNCH_CHECK(ctx);
	return TW_OK;
}
