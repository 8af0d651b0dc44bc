#include <cstdio>
#include <cuda_runtime.h>
template<int MODE> __global__ void k(float *out, int iters, float a, float b) {
	float2 x0 = make_float2(threadIdx.x*1e-3f, 1.0f), x1 = make_float2(2.0f, 3.0f), x2 = make_float2(0.5f, 0.25f), x3 = make_float2(1.5f, 2.5f);
	float2 const m = make_float2(a, a), c = make_float2(b, b);
	for (int i = 0; i < iters; ++i) {
		if (MODE == 0) { // scalar mul + add: 16 instr for 16 flops... (8 mul, 8 add)
			x0.x = __fadd_rn(__fmul_rn(x0.x, m.x), c.x); x0.y = __fadd_rn(__fmul_rn(x0.y, m.y), c.y);
			x1.x = __fadd_rn(__fmul_rn(x1.x, m.x), c.x); x1.y = __fadd_rn(__fmul_rn(x1.y, m.y), c.y);
			x2.x = __fadd_rn(__fmul_rn(x2.x, m.x), c.x); x2.y = __fadd_rn(__fmul_rn(x2.y, m.y), c.y);
			x3.x = __fadd_rn(__fmul_rn(x3.x, m.x), c.x); x3.y = __fadd_rn(__fmul_rn(x3.y, m.y), c.y);
		} else { // packed
			x0 = __fadd2_rn(__fmul2_rn(x0, m), c); x1 = __fadd2_rn(__fmul2_rn(x1, m), c);
			x2 = __fadd2_rn(__fmul2_rn(x2, m), c); x3 = __fadd2_rn(__fmul2_rn(x3, m), c);
		}
	}
	out[blockIdx.x*blockDim.x + threadIdx.x] = x0.x + x0.y + x1.x + x1.y + x2.x + x2.y + x3.x + x3.y;
}
int main() {
	float *d; cudaMalloc(&d, 148*16*256*sizeof(float));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	int const iters = 20000;
	for (int mode = 0; mode < 2; ++mode) {
		for (int rep = 0; rep < 3; ++rep) {
			cudaEventRecord(e0);
			if (mode == 0) k<0><<<148*16, 256>>>(d, iters, 0.999f, 0.001f); else k<1><<<148*16, 256>>>(d, iters, 0.999f, 0.001f);
			cudaEventRecord(e1); cudaEventSynchronize(e1);
			float ms; cudaEventElapsedTime(&ms, e0, e1);
			double const flops = 16.0*iters*148*16*256;
			printf("mode %d: %.3f ms  %.2f Tflop/s (mul+add counted separately)\n", mode, ms, flops/ms/1e9);
		}
	}
	return 0;
}
