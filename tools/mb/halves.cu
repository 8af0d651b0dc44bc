// Does the ORDER of packed (FMUL2/FFMA2: both 16-lane halves of the FP32 pipe for 2 cycles) and scalar (FMUL/FADD: one half for 2 cycles) instructions matter?
// Same instruction counts, different interleavings; independent registers; 1..6 warps per sub-partition. Prints cycles per 16-instruction group per sub-partition
// (ideal: 8 packed x 2 + 8 scalar x 1 = 24).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb/halves tools/mb/halves.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define P(i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(m))
#define S(i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a))
#define A(i) asm volatile("add.s32 %0, %0, %1;" : "+r"(n[i]) : "r"(k))
template<int PAT> __global__ void __launch_bounds__(128, 8) k(float *out, int iters, float a, int kk, long long *cycles) {
	u64 p[8], m; float x[8]; int n[8]; int const k = kk;
	for (int i = 0; i < 8; ++i) {float const v = 1.0f + threadIdx.x*1e-4f + i*1e-3f; asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(v), "f"(v + 0.5f)); x[i] = v; n[i] = threadIdx.x + i;}
	asm("mov.b64 %0, {%1, %2};" : "=l"(m) : "f"(a), "f"(a));
	long long const t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < iters; ++it) {
		if (PAT == 0) {P(0); S(0); P(1); S(1); P(2); S(2); P(3); S(3); P(4); S(4); P(5); S(5); P(6); S(6); P(7); S(7);}
		if (PAT == 1) {P(0); P(1); S(0); S(1); P(2); P(3); S(2); S(3); P(4); P(5); S(4); S(5); P(6); P(7); S(6); S(7);}
		if (PAT == 2) {P(0); P(1); P(2); P(3); P(4); P(5); P(6); P(7); S(0); S(1); S(2); S(3); S(4); S(5); S(6); S(7);}
		if (PAT == 3) {P(0); P(1); P(2); P(3); P(4); P(5); P(6); P(7); P(0); P(1); P(2); P(3); P(4); P(5); P(6); P(7);}   // packed only: 32
		if (PAT == 4) {S(0); S(1); S(2); S(3); S(4); S(5); S(6); S(7); S(0); S(1); S(2); S(3); S(4); S(5); S(6); S(7);}   // scalar only: 16
		if (PAT == 5) {P(0); S(0); S(1); P(1); S(2); S(3); P(2); S(4); S(5); P(3); S(6); S(7); P(4); S(0); S(1); P(5);}   // 6 P + 10 S, scalars in pairs: 22
		if (PAT == 6) {P(0); S(0); P(1); S(1); P(2); S(2); P(3); S(3); P(4); S(4); P(5); S(5); S(6); S(7); S(0); S(1);}   // 6 P + 10 S, singles first: 22
		if (PAT == 7) {P(0); A(0); P(1); A(1); P(2); A(2); P(3); A(3); P(4); A(4); P(5); A(5); P(6); A(6); P(7); A(7);}   // packed + integer add (ALU pipe): 16 if hidden
		if (PAT == 8) {P(0); A(0); S(0); P(1); A(1); S(1); P(2); A(2); S(2); P(3); A(3); S(3); P(4); S(4); P(5); S(5);}   // 6 P, 6 S, 4 A
	}
	long long const t1 = clock64();
	float s = 0; for (int i = 0; i < 8; ++i) {s += x[i] + n[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));}
	out[blockIdx.x*blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template<int PAT> void run(float *d, long long *dc, const char *name, int ideal) {
	for (int bps = 1; bps <= 6; bps += (bps == 1) ? 1 : 2) { // blocks of 4 warps per SM = warps per sub-partition
		int const nb = 148*bps, iters = 4000;
		static long long hc[148*8];
		for (int rep = 0; rep < 2; ++rep) {k<PAT><<<nb, 128>>>(d, iters, 0.9999f, 3, dc); cudaDeviceSynchronize();}
		cudaMemcpy(hc, dc, nb*sizeof(long long), cudaMemcpyDeviceToHost);
		double cyc = 0; for (int i = 0; i < nb; ++i) cyc += hc[i]; cyc /= nb;
		printf("%-46s %d warps/SMSP: %6.2f clk per group per SMSP (ideal %d)\n", name, bps, cyc/iters/bps, ideal);
	}
}
int main() {
	float *d; long long *dc; cudaMalloc(&d, 148*8*128*sizeof(float)); cudaMalloc(&dc, 148*8*sizeof(long long));
	run<0>(d, dc, "P S P S ... (8 P + 8 S alternating)", 24); run<1>(d, dc, "P P S S ... (pairs)", 24); run<2>(d, dc, "8 P then 8 S", 24);
	run<3>(d, dc, "16 P", 32); run<4>(d, dc, "16 S", 16); run<5>(d, dc, "6 P + 10 S, scalars in pairs", 22); run<6>(d, dc, "6 P + 10 S, scalars single between P", 22);
	run<7>(d, dc, "8 P + 8 IADD alternating", 16); run<8>(d, dc, "6 P + 6 S + 4 IADD", 18);
	cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) {printf("error %s\n", cudaGetErrorString(e)); return 1;}
	return 0;
}
