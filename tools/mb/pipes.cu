// Issue/pipe rules of the B200 SM sub-partition that bound the noise kernel: how many cycles a packed fp32x2 instruction holds the FMA pipe for each operand form,
// and whether instructions of other pipes (ALU integer add / min-max, LDS, FRND on the XU pipe) issue in the shadow of a packed instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb/pipes tools/mb/pipes.cu && tools/mb/pipes
// Output: warp instructions per cycle per sub-partition for every mix (1.0 = one instruction issued every cycle).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__constant__ float2 C_ONE = {1.0f, 1.0f};
#define FMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define MUL2(d, a, b)    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define FMA1(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define MUL1(d, a, b)    asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define ADD1(d, a, b)    asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define IADD(d, a, b)    asm volatile("add.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define FMAX(d, a, b)    asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define FLOOR(d, a)      asm volatile("cvt.rmi.f32.f32 %0, %1;" : "=f"(d) : "f"(a))
#define LDS(d, a)        asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(d) : "r"(a))

enum {M_FMA2_3REG, M_FMA2_CONST, M_MUL2, M_FMA1, M_MUL1, M_ADD1, M_FMA2_IADD, M_FMA2_2IADD, M_FMA2_FMAX, M_FMA2_LDS, M_FMA2_FLOOR, M_FMA1_IADD, M_IADD, M_MUL2_IADD, M_FMA2C_IADD, M_MIX_KERNEL, M_FMA2_DENORM, M_COUNT};
const char *NAMES[] = {"FFMA2 r,r,r", "FFMA2 r,c[],r", "FMUL2 r,r", "FFMA r,r,r", "FMUL r,r", "FADD r,r", "FFMA2 + IADD", "FFMA2 + 2 IADD", "FFMA2 + FMNMX", "FFMA2 + LDS (1 per 2)", "FFMA2 + FRND (1 per 4)",
                       "FFMA + IADD", "IADD only", "FMUL2 + IADD", "FFMA2 r,c[],r + IADD", "kernel mix: 23 FFMA2c,15 FMUL2,12 FMUL,5 FADD,20 IADD,8 LDS,5 FRND,8 FMNMX", "FFMA2 r,r,r with denormal multiplier/addend/result (table addressing without IADD)"};
const int INSTR_PER_ITER[] = {8, 8, 8, 8, 8, 8, 16, 24, 16, 12, 10, 16, 8, 16, 16, 96, 8};

template<int MODE> __global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b, long long *cycles) {
	__shared__ float sm[1024];
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i*1e-3f;
	__syncthreads();
	u64 p[8], m[8], c[8];
	float x[8], y[8], z[8], w[8], l[8], f[8], g[8];
	u64 q[8];
	int n[8];
	unsigned addr = (unsigned)__cvta_generic_to_shared(sm) + 4*threadIdx.x;
	for (int i = 0; i < 8; ++i) {
		float const v = threadIdx.x*1e-3f + i;
		asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(v), "f"(v + 0.5f));
		asm("mov.b64 %0, {%1, %2};" : "=l"(m[i]) : "f"(a + i*1e-6f), "f"(a - i*1e-6f));
		asm("mov.b64 %0, {%1, %2};" : "=l"(c[i]) : "f"(b + i*1e-6f), "f"(b - i*1e-6f));
		x[i] = v; y[i] = a + i*1e-6f; n[i] = threadIdx.x + i; z[i] = v + 1.0f; w[i] = v + 2.0f; l[i] = 0.0f; f[i] = v*3.7f; g[i] = v; q[i] = p[i];
	}
	u64 one; asm("mov.b64 %0, {%1, %2};" : "=l"(one) : "f"(C_ONE.x), "f"(C_ONE.y));
	u64 den, dc[8];
	asm("mov.b64 %0, {%1, %2};" : "=l"(den) : "f"(__uint_as_float(128u)), "f"(__uint_as_float(128u)));
	for (int i = 0; i < 8; ++i) {asm("mov.b64 %0, {%1, %2};" : "=l"(dc[i]) : "f"(__uint_as_float(16u*(threadIdx.x & 7) + i)), "f"(__uint_as_float(16u*(threadIdx.x & 7) + 4u*i)));}
	if (MODE == M_FMA2_DENORM) {for (int i = 0; i < 8; ++i) {asm("mov.b64 %0, {%1, %2};" : "=l"(m[i]) : "f"((float)(i & 3)), "f"((float)(i & 5)));}}
	long long const t0 = clock64();
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			if (MODE == M_FMA2_3REG || MODE == M_FMA2_IADD || MODE == M_FMA2_2IADD || MODE == M_FMA2_FMAX || MODE == M_FMA2_LDS || MODE == M_FMA2_FLOOR) FMA2(p[i], p[i], m[i], c[i]);
			if (MODE == M_FMA2_CONST || MODE == M_FMA2C_IADD) FMA2(p[i], p[i], one, c[i]);
			if (MODE == M_FMA2_DENORM) FMA2(dc[i], m[i], den, dc[i]); // m = small integers, den = 128*2^-149, dc = denormal addend: result bits = an integer
			if (MODE == M_MUL2 || MODE == M_MUL2_IADD) MUL2(p[i], p[i], m[i]);
			if (MODE == M_FMA1 || MODE == M_FMA1_IADD) FMA1(x[i], x[i], y[i], y[(i + 1) & 7]);
			if (MODE == M_MUL1) MUL1(x[i], x[i], y[i]);
			if (MODE == M_ADD1) ADD1(x[i], x[i], y[i]);
			if (MODE == M_FMA2_IADD || MODE == M_FMA1_IADD || MODE == M_IADD || MODE == M_MUL2_IADD || MODE == M_FMA2C_IADD || MODE == M_FMA2_2IADD) IADD(n[i], n[i], n[(i + 1) & 7]);
			if (MODE == M_FMA2_2IADD) IADD(n[i], n[i], n[(i + 3) & 7]);
			if (MODE == M_FMA2_FMAX) FMAX(x[i], x[i], y[i]);
			if (MODE == M_FMA2_LDS && (i & 1)) LDS(x[i], addr);
			if (MODE == M_FMA2_FLOOR && (i & 3) == 0) FLOOR(x[i], y[i]);
		}
		if (MODE == M_MIX_KERNEL) { // per-evaluation instruction mix of noise_grid2_kernel (profiles/ncu_noise_grid2_kernel_r02.json), independent chains
#include "pipes_mix.inc"
		}
	}
	long long const t1 = clock64();
	float s = 0;
	for (int i = 0; i < 8; ++i) {s += x[i] + y[i] + n[i] + z[i] + w[i] + l[i] + f[i] + g[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)q[i]) + __uint_as_float((unsigned)(q[i] >> 32)) + (float)(unsigned)dc[i] + (float)(unsigned)(dc[i] >> 32);}
	out[blockIdx.x*blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template<int MODE> void run(float *d, long long *dc, int iters, int blocks_per_sm) {
	int const nb = 148*blocks_per_sm;
	long long hc[148*8];
	for (int rep = 0; rep < 2; ++rep) {
		k<MODE><<<nb, 256>>>(d, iters, 0.999f, 0.001f, dc);
		cudaDeviceSynchronize();
	}
	cudaMemcpy(hc, dc, nb*sizeof(long long), cudaMemcpyDeviceToHost);
	double cyc = 0; for (int i = 0; i < nb; ++i) cyc += hc[i]; cyc /= nb;
	int const instr = (MODE == M_MIX_KERNEL) ? 96 : INSTR_PER_ITER[MODE];
	double const warps_per_smsp = blocks_per_sm*8/4.0;
	printf("%-90s %d warps/SMSP: %.3f instr/clk/SMSP  (%.2f clk per %d-instr group per warp)\n", NAMES[MODE], (int)warps_per_smsp, instr*(double)iters*warps_per_smsp/cyc, cyc/iters, instr);
}
template<int MODE> void run_all(float *d, long long *dc) {run<MODE>(d, dc, 4000, 1); run<MODE>(d, dc, 4000, 3); run<MODE>(d, dc, 4000, 4);}
int main() {
	float *d; long long *dc; cudaMalloc(&d, 148*8*256*sizeof(float)); cudaMalloc(&dc, 148*8*sizeof(long long));
	run_all<M_FMA2_3REG>(d, dc); run_all<M_FMA2_CONST>(d, dc); run_all<M_MUL2>(d, dc); run_all<M_FMA1>(d, dc); run_all<M_MUL1>(d, dc); run_all<M_ADD1>(d, dc);
	run_all<M_IADD>(d, dc); run_all<M_FMA2_IADD>(d, dc); run_all<M_FMA2_2IADD>(d, dc); run_all<M_FMA2C_IADD>(d, dc); run_all<M_MUL2_IADD>(d, dc); run_all<M_FMA1_IADD>(d, dc);
	run_all<M_FMA2_FMAX>(d, dc); run_all<M_FMA2_LDS>(d, dc); run_all<M_FMA2_FLOOR>(d, dc); run_all<M_MIX_KERNEL>(d, dc); run_all<M_FMA2_DENORM>(d, dc);
	cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) {printf("error %s\n", cudaGetErrorString(e)); return 1;}
	return 0;
}
