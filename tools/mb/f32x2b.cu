#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) {u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;}
// MODE 0: scalar FMUL+FADD (unfused), 1: packed FMUL2+FADD2 (unfused), 2: scalar FFMA, 3: packed FFMA2
template<int MODE> __global__ void k(float *out, int iters, float a, float b) {
	float x[8]; for (int i = 0; i < 8; ++i) x[i] = threadIdx.x*1e-3f + i;
	u64 p[4]; for (int i = 0; i < 4; ++i) p[i] = pk(x[2*i], x[2*i+1]);
	u64 const m = pk(a, a), c = pk(b, b);
	for (int it = 0; it < iters; ++it) {
		if (MODE == 0) {for (int i = 0; i < 8; ++i) x[i] = __fadd_rn(__fmul_rn(x[i], a), b);}
		if (MODE == 1) {for (int i = 0; i < 4; ++i) p[i] = add2(mul2(p[i], m), c);}
		if (MODE == 2) {for (int i = 0; i < 8; ++i) x[i] = __fmaf_rn(x[i], a, b);}
		if (MODE == 3) {for (int i = 0; i < 4; ++i) p[i] = fma2(p[i], m, c);}
	}
	float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
	for (int i = 0; i < 4; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
	out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}
template<int MODE> void run(float *d, int iters) {
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int rep = 0; rep < 2; ++rep) {
		cudaEventRecord(e0); k<MODE><<<148*16, 256>>>(d, iters, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		double const elem_ops = 8.0*iters*148*16*256; // 8 elements updated per thread per iteration
		if (rep) printf("mode %d: %.3f ms  %.2f T element-updates/s\n", MODE, ms, elem_ops/ms/1e9);
	}
}
int main() {
	float *d; cudaMalloc(&d, 148*16*256*sizeof(float));
	run<0>(d, 20000); run<1>(d, 20000); run<2>(d, 20000); run<3>(d, 20000);
	return 0;
}
