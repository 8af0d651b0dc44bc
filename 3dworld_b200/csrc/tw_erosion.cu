// tw_erosion.cu - droplet hydraulic erosion (sm_100a). Replaces apply_erosion() (src/erosion.cpp:14-164).
//
// The reference algorithm is a Lagrangian droplet walk: droplet k sees every height written by droplets < k, so one heightmap is a
// serial dependency chain (SURVEY.md section 7 "erosion order dependence"). The parallelism the reference's callers expose is ACROSS
// heightmaps: tile_t::create_zvals erodes every tile independently with the same droplet seeds (src/tiled_mesh.cpp:515). Mapping:
//   pad_kernel      (:31-37)   clamped PAD=4 border copy, fully parallel, coalesced
//   droplet_kernel  (:66-155)  ONE WARP PER HEIGHTMAP, droplets in the reference's serial order. The scalar droplet state is kept
//                              redundantly in all 32 lanes (uniform control flow, no shuffles); the 2x2 bilinear deposit is done by
//                              lanes 0-3 and the 4x4 erode brush by lanes 0-15, one cell each, so a brush costs 4 row-coalesced
//                              read-modify-writes instead of 16 serial ones. Heights live in global memory and are served from L2/L1.
//   unpad_kernel    (:158-162) remove border, clamp to min_zval
// Bit-exactness: IEEE sqrt/div, no FMA contraction (-fmad=false), std::min/max argument order preserved (NaN semantics),
// the random-direction fallback (:84-87) reads cos/sin from a 1e6-entry table built with the HOST libm (rand_float() has only 1e6 values).
#include "tw_internal.h"
#include <float.h>
#include <stdlib.h>

namespace {

constexpr int PAD = 4;

__device__ __forceinline__ float smin(float a, float b) {return (b < a) ? b : a;} // std::min
__device__ __forceinline__ float smax(float a, float b) {return (a < b) ? b : a;} // std::max
__device__ __forceinline__ int clampi(int v, int hi) {return max(min(v, hi), 0);}

struct EParams {
	float erode_amount, wpz_minus_half_dxy, zmin, zrange, relh_adj_tex, clip_hd1;
};

// Also counts, per heightmap, the cells above the ocean-stop level (src/erosion.cpp:98): droplets that start below it die in one move, the
// others walk downhill, so this count predicts the heightmap's total droplet work (correlation 0.95 on the BASELINE terrain) and is used to
// schedule the heaviest heightmaps first (the work per heightmap is heavy-tailed: median 3, mean 26, max > 140 moves per droplet).
__global__ void pad_kernel(const float *__restrict__ in, float *__restrict__ out, int xsize, int ysize, int NX, int NY, float work_level, unsigned *__restrict__ work) {
	int const x = blockIdx.x*blockDim.x + threadIdx.x, y = blockIdx.y;
	size_t const tile = blockIdx.z;
	bool above = false;
	if (x < NX) {
		int const sx = clampi(x - PAD, xsize - 1), sy = clampi(y - PAD, ysize - 1);
		float const v = __ldg(in + tile*xsize*ysize + (size_t)sy*xsize + sx);
		out[tile*NX*NY + (size_t)y*NX + x] = v;
		above = !(v < work_level);
	}
	unsigned const n = __popc(__ballot_sync(0xffffffffu, above));
	if (work && n && (threadIdx.x & 31) == 0) {atomicAdd(work + tile, n);}
}

// counting sort of the heightmap indices by descending work estimate (256 bins; order within a bin is irrelevant: heightmaps are independent)
constexpr int WORK_BINS = 256;
__global__ void order_hist_kernel(const unsigned *__restrict__ work, unsigned nt, unsigned max_work, unsigned *__restrict__ hist) {
	unsigned const t = blockIdx.x*blockDim.x + threadIdx.x;
	if (t >= nt) return;
	unsigned const bin = (WORK_BINS - 1) - min((unsigned)(WORK_BINS - 1), (unsigned)(((unsigned long long)work[t]*(WORK_BINS - 1))/max_work));
	atomicAdd(hist + bin, 1u);
}
__global__ void order_scan_kernel(unsigned *__restrict__ hist) { // exclusive prefix sum of 256 bins, one warp
	unsigned const lane = threadIdx.x;
	unsigned v[WORK_BINS/32], sum = 0;
	for (int i = 0; i < WORK_BINS/32; ++i) {v[i] = hist[lane*(WORK_BINS/32) + i]; sum += v[i];}
	unsigned incl = sum;
	for (int o = 1; o < 32; o <<= 1) {unsigned const n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += n;}
	unsigned run = incl - sum;
	for (int i = 0; i < WORK_BINS/32; ++i) {hist[lane*(WORK_BINS/32) + i] = run; run += v[i];}
}
__global__ void order_scatter_kernel(const unsigned *__restrict__ work, unsigned nt, unsigned max_work, unsigned *__restrict__ cursor, unsigned *__restrict__ order) {
	unsigned const t = blockIdx.x*blockDim.x + threadIdx.x;
	if (t >= nt) return;
	unsigned const bin = (WORK_BINS - 1) - min((unsigned)(WORK_BINS - 1), (unsigned)(((unsigned long long)work[t]*(WORK_BINS - 1))/max_work));
	order[atomicAdd(cursor + bin, 1u)] = t;
}

__global__ void unpad_kernel(const float *__restrict__ padded, float *__restrict__ out, int xsize, int ysize, int NX, int NY,
	const float *__restrict__ min_zvals, float min_zval_all)
{
	int const x = blockIdx.x*blockDim.x + threadIdx.x, y = blockIdx.y;
	size_t const tile = blockIdx.z;
	if (x >= xsize) return;
	float const mz = min_zvals ? __ldg(min_zvals + tile) : min_zval_all;
	out[tile*xsize*ysize + (size_t)y*xsize + x] = smax(mz, padded[tile*NX*NY + (size_t)(y + PAD)*NX + x + PAD]);
}

// rand_gen_t core (src/rand_gen.h:22-26) in 32-bit: all intermediates fit (Schrage factorisation), states stay in [0, 2^31)
struct Rng {
	int s1, s2;
	__device__ __forceinline__ int rand() {
		if ((s1 = 40014*(s1%53668) - 12211*(s1/53668)) < 0) s1 += 2147483563;
		if ((s2 = 40692*(s2%52774) - 3791 *(s2/52774)) < 0) s2 += 2147483399;
		int r = s1 - s2;
		if (r < 1) r += 2147483562;
		return r;
	}
};

// One heightmap per group of G lanes (G = 1, 2, 4, 8, 16, 32; 32/G heightmaps per warp). The droplet's scalar state is replicated in the
// G lanes of its group (group-uniform control flow, no shuffles); the 4 deposit corners and the 16 brush cells are dealt round-robin to
// the lanes of the group (c = sub, sub+G, ...: ascending c is the reference's z-outer/x-inner order, so G == 1 is literally the serial
// loop). The droplet loop is flattened into one state machine per group (init-droplet / step) so that groups whose droplets end at
// different times stay converged at the top of the loop. G trades redundant ALU work (G = 32: every lane repeats the ~250-instruction
// step for ONE map) against memory coalescing and in-flight parallelism (G = 1: no redundancy, but 32 unrelated maps per load
// instruction and 32x more maps needed to fill the machine); twi_erode picks G from the number of heightmaps.
// SHARED (tw_erode_parallel): the reference's multi-threaded mode, `#pragma omp parallel for schedule(dynamic,1)` over the droplets of ONE
// heightmap (src/erosion.cpp:66): `ntiles` groups play the OpenMP threads, each takes the next droplet index from an atomic counter (the
// dynamic,1 schedule), reads bypass L1 (the CPU's caches are coherent) and the read-modify-writes are float atomics (the reference's are
// unsynchronised). One group => exactly the serial order.
template<int G, bool SHARED = false>
__global__ void __launch_bounds__(128)
droplet_kernel(float *__restrict__ padded, unsigned ntiles, int xsize, int ysize, unsigned num_iters, EParams E,
	const float2 *__restrict__ dir_table, unsigned long long *__restrict__ steps_out, const unsigned *__restrict__ order, unsigned *__restrict__ next_droplet = nullptr)
{
	constexpr int TPW = 32/G; // heightmaps per warp
	int const lane = threadIdx.x & 31, sub = lane % G, grp = lane / G;
	unsigned const warp = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	unsigned const slot = warp*TPW + grp;                                       // position in the heaviest-first schedule
	unsigned const tile = (slot < ntiles) ? (order ? __ldg(order + slot) : slot) : ntiles;
	unsigned const gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (grp*G));
	int const NX = xsize + 2*PAD, NY = ysize + 2*PAD;
	float *mh = padded + (SHARED ? (size_t)0 : (size_t)((tile < ntiles) ? tile : 0)*NX*NY);
	float const Kq=10, Kw=0.001f, Kr=0.9f, Kd=0.02f, Ki=0.1f, minSlope=0.05f, g=20, Kg=g*2;
	unsigned const MAX_PATH_LEN = 4u*(unsigned)NX*(unsigned)NY;
	float const erode_amount = E.erode_amount;
	unsigned long long steps = 0;
	bool active = (tile < ntiles), in_droplet = false;
	unsigned iter = 0, numMoves = 0;
	Rng rgen; rgen.s1 = rgen.s2 = 1;
	int xi = 0, zi = 0;
	float xp=0, zp=0, xf=0, zf=0, s=0, v=0, w=1, dx=0, dz=0, h=0, h00=0, h10=0, h01=0, h11=0;

#define HMAP(x, y) mh[(size_t)NX*clampi((y), NY-1) + clampi((x), NX-1)]
#define HLOAD(ptr)       (SHARED ? __ldcg(ptr) : *(ptr))
#define HADD(ptr, delta) {if (SHARED) {atomicAdd((ptr), (delta));} else {*(ptr) += (delta);}}
	// DEPOSIT(H): src/erosion.cpp:42-54; corner c of the 2x2 cell goes to lane c % G (inside cells are distinct => no aliasing between lanes)
#define DEPOSIT(H) { \
	_Pragma("unroll") \
	for (int c = sub; c < 4; c += G) { \
		int const X = xi + (c & 1), Z = zi + (c >> 1); \
		float const W = ((c & 1) ? xf : (1-xf))*((c >> 1) ? zf : (1-zf)); \
		float const delta = ds*erode_amount*W; \
		if ((unsigned)X < (unsigned)NX && (unsigned)Z < (unsigned)NY) {HADD(&mh[NX*Z + X], delta)} \
	} \
	if (G > 1) {__syncwarp(gmask);} \
	(H) += ds; }

	for (;;) {
		if (active && !in_droplet) { // next droplet of this group's heightmap (src/erosion.cpp:67-73)
			if (SHARED) { // schedule(dynamic,1): the group's leader draws the next droplet
				unsigned nd = 0;
				if (sub == 0) {nd = atomicAdd(next_droplet, 1u);}
				iter = __shfl_sync(gmask, nd, grp*G);
			}
			if (iter >= num_iters) {active = false;}
			else {
				rgen.s1 = (int)iter + 11; rgen.s2 = 79*(int)iter + 121;
				xi = PAD + (rgen.rand()%xsize);
				zi = PAD + (rgen.rand()%ysize);
				xp=xi; zp=zi; xf=0; zf=0; s=0; v=0; w=1; dx=0; dz=0;
				h=HLOAD(&HMAP(xi, zi)); h00=h; h10=HLOAD(&HMAP(xi+1, zi)); h01=HLOAD(&HMAP(xi, zi+1)); h11=HLOAD(&HMAP(xi+1, zi+1));
				numMoves = 0; in_droplet = true; ++iter;
			}
		}
		if (!__any_sync(0xffffffffu, active)) break;
		if (!active) continue;
		if (numMoves >= MAX_PATH_LEN) {in_droplet = false; continue;} // "droplet path is too long" (src/erosion.cpp:153)
		++numMoves; ++steps;
		{ // ---- one move of the droplet (src/erosion.cpp:76-152) ----
			float const gx=h00+h01-h10-h11, gz=h00+h10-h01-h11;
			dx=(dx-gx)*Ki+gx;
			dz=(dz-gz)*Ki+gz;
			float const dl=__fsqrt_rn(dx*dx+dz*dz);
			if (dl<=FLT_EPSILON) { // pick random dir: a = rand_float()*TWO_PI, rand_float() = 1e-6*(rand()%1000000)
				float2 const cs = __ldg(dir_table + (rgen.rand()%1000000));
				dx=cs.x; dz=cs.y;
			}
			else {dx=__fdiv_rn(dx, dl); dz=__fdiv_rn(dz, dl);}
			float const nxp=xp+dx, nzp=zp+dz;
			int nxi=__float2int_rd(nxp), nzi=__float2int_rd(nzp); // (int)floor(.)
			if (!(fabsf(nxp) < 2147483648.0f && fabsf(nzp) < 2147483648.0f)) { // NaN / out of int range: x86 cvttss2si yields INT_MIN -> "outside" next step
				nxi=tw_x86_f2i(floorf(nxp)); nzi=tw_x86_f2i(floorf(nzp));
			}
			float const nxf=nxp-(float)nxi, nzf=nzp-(float)nzi;
			float nh00, nh10, nh01, nh11;
			if ((unsigned)nxi < (unsigned)(NX-1) && (unsigned)nzi < (unsigned)(NY-1)) { // common case: no clamping needed
				float const *q = mh + (nzi*NX + nxi);
				nh00 = HLOAD(q); nh10 = HLOAD(q + 1); q += NX; nh01 = HLOAD(q); nh11 = HLOAD(q + 1);
			}
			else {nh00=HLOAD(&HMAP(nxi, nzi)); nh10=HLOAD(&HMAP(nxi+1, nzi)); nh01=HLOAD(&HMAP(nxi, nzi+1)); nh11=HLOAD(&HMAP(nxi+1, nzi+1));}
			float const nh=(nh00*(1-nxf)+nh10*nxf)*(1-nzf)+(nh01*(1-nxf)+nh11*nxf)*nzf;
			if (smax(smax(nh00, nh10), smax(nh01, nh11)) < E.wpz_minus_half_dxy) {in_droplet = false; continue;} // reached ocean water

			bool const outside = ((unsigned)xi >= (unsigned)NX || (unsigned)zi >= (unsigned)NY);
			if (nh>=h || outside) {
				float ds=(nh-h)+0.001f;
				if (ds>=s || outside) {
					ds=s;
					DEPOSIT(h)
					s=0;
					in_droplet = false; continue;
				}
				DEPOSIT(h)
				s-=ds;
				v=0;
			}
			float dh=h-nh;
			float const q=smax(dh, minSlope)*v*w*Kq;
			float ds=s-q;
			if (ds>=0) { // deposit
				ds*=Kd;
				DEPOSIT(dh)
				s-=ds;
			}
			else { // erode
				ds*=-Kr;
				ds=smin(ds, dh*0.99f);
				{ // get_bare_ls_tid(nh) == ROCK_TEX ? 0.5 : 2.0 (src/Textures.cpp:1284-1287); x0.5 / x2 are exact in fp32
					float const relh = E.relh_adj_tex + __fdiv_rn(nh - E.zmin, E.zrange);
					ds *= (relh > E.clip_hd1) ? 0.5f : 2.0f;
				}
				bool const interior = ((unsigned)(xi - 1) < (unsigned)(NX - 3) && (unsigned)(zi - 1) < (unsigned)(NY - 3));
				if (interior || G == 1) { // 16 distinct cells dealt to the lanes of the group (G == 1: the reference's serial loop, clamped)
#pragma unroll
					for (int c = sub; c < 16; c += G) {
						int const x = xi + (c & 3) - 1, z = zi + (c >> 2) - 1;
						float const zo=(float)z-zp, zo2=zo*zo, xo=(float)x-xp;
						float wgt=1-(xo*xo+zo2)*0.25f;
						if (!(wgt<=0)) {
							wgt*=0.1591549430918953f;
							float const delta=ds*erode_amount*wgt;
							if (interior) {HADD(&mh[NX*z + x], -delta)} else {HADD(&HMAP(x, z), -delta)}
						}
					}
				}
				else if (sub == 0) { // border: clamped indices may alias, keep the reference's serial order
					for (int z=zi-1; z<=zi+2; ++z) {
						float const zo=(float)z-zp, zo2=zo*zo;
						for (int x=xi-1; x<=xi+2; ++x) {
							float const xo=(float)x-xp;
							float wgt=1-(xo*xo+zo2)*0.25f;
							if (wgt<=0) continue;
							wgt*=0.1591549430918953f;
							float const delta=ds*erode_amount*wgt;
							HADD(&HMAP(x, z), -delta)
						}
					}
				}
				if (G > 1) {__syncwarp(gmask);}
				dh-=ds;
				s+=ds;
			}
			v=__fsqrt_rn(v*v+Kg*dh);
			w*=1-Kw;
			xp=nxp; zp=nzp; xi=nxi; zi=nzi; xf=nxf; zf=nzf;
			h=nh; h00=nh00; h10=nh10; h01=nh01; h11=nh11;
		}
	}
#undef HMAP
#undef HLOAD
#undef HADD
#undef DEPOSIT
	if (sub == 0 && steps_out && steps) {atomicAdd(steps_out, steps);}
}

template<int G>
void launch_droplets(cudaStream_t st, float *d_pad, unsigned nt, int xsize, int ysize, unsigned num_iters, EParams const &E, const float2 *dir, unsigned long long *d_steps, const unsigned *order) {
	unsigned const warps_per_block = 2, tiles_per_block = warps_per_block*(32/G);
	droplet_kernel<G><<<(nt + tiles_per_block - 1)/tiles_per_block, 32*warps_per_block, 0, st>>>(d_pad, nt, xsize, ysize, num_iters, E, dir, d_steps, order);
}

template<int G>
void launch_droplets_shared(cudaStream_t st, float *d_pad, unsigned ngroups, int xsize, int ysize, unsigned num_iters, EParams const &E, const float2 *dir, unsigned long long *d_steps, unsigned *d_next) {
	unsigned const warps_per_block = 2, groups_per_block = warps_per_block*(32/G);
	droplet_kernel<G, true><<<(ngroups + groups_per_block - 1)/groups_per_block, 32*warps_per_block, 0, st>>>(d_pad, ngroups, xsize, ysize, num_iters, E, dir, d_steps, nullptr, d_next);
}

// lanes per heightmap: the smallest group that still gives ~12 warps per SM (148 SMs), see the kernel comment
int pick_group(unsigned ntiles) {
	const char *env = getenv("TW_EROSION_LANES");
	if (env) {int const g = atoi(env); if (g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32) return g;}
	// measured on B200 (tools/bench_erosion.py, 258^2 tiles, heaviest-first schedule): 16384 maps: G=32 0.18 s, G=16 0.23 s, G=8 0.27 s;
	// 65536 maps: G=32 0.53 s, G=16 0.40 s, G=8 0.40 s. With >= ~16k warps the kernel is issue-bound (78 % issue slots at G=32), so sharing a
	// warp between maps pays; below that it is latency/tail-bound and one warp per map is fastest. => smallest G >= 8 that keeps 16384 warps.
	for (int g = 8; g < 32; g *= 2) {if ((unsigned long long)ntiles*g >= 32ull*16384ull) return g;}
	return 32;
}

} // namespace

static EParams make_eparams(const tw_erosion_params *p) {
	EParams E;
	E.erode_amount = p->erode_amount;
	E.wpz_minus_half_dxy = p->water_plane_z - p->half_dxy;
	E.zmin = p->zmin; E.zrange = p->zmax - p->zmin;
	E.relh_adj_tex = p->relh_adj_tex; E.clip_hd1 = p->clip_hd1;
	return E;
}

size_t twi_erode_scratch_bytes(uint32_t chunk, int xsize, int ysize) {
	size_t const padded_elems = (size_t)(xsize + 2*PAD)*(ysize + 2*PAD);
	size_t const pad_bytes = ((size_t)chunk*padded_elems*sizeof(float) + 255) & ~(size_t)255;
	return pad_bytes + (((size_t)chunk*2 + WORK_BINS)*sizeof(unsigned) + 255 & ~(size_t)255);
}

// Enqueue pad -> schedule -> droplets -> unpad for nt <= 65535 heightmaps on `st`, using `scratch` (twi_erode_scratch_bytes(capacity,..) bytes).
// No synchronisation; d_steps (device counter) accumulates the droplet moves.
int twi_erode_enqueue(tw_ctx *ctx, cudaStream_t st, void *scratch, uint32_t capacity, float *maps, uint32_t nt, int xsize, int ysize,
                      const float *d_min_zvals, float min_zval_all, uint32_t num_iters, const tw_erosion_params *p, unsigned long long *d_steps)
{
	int const NX = xsize + 2*PAD, NY = ysize + 2*PAD;
	size_t const padded_elems = (size_t)NX*NY;
	EParams const E = make_eparams(p);
	size_t const pad_bytes = ((size_t)capacity*padded_elems*sizeof(float) + 255) & ~(size_t)255;
	float *d_pad = (float *)scratch;
	unsigned *d_work = (unsigned *)((char *)scratch + pad_bytes), *d_hist = d_work + capacity, *d_order = d_hist + WORK_BINS;
	bool const schedule = (nt > 148u*4u); // with few heightmaps everything is resident at once anyway
	if (schedule) {TW_CUDA(ctx, cudaMemsetAsync(d_work, 0, ((size_t)capacity + WORK_BINS)*sizeof(unsigned), st));}
	pad_kernel<<<dim3((NX + 255)/256, NY, nt), 256, 0, st>>>(maps, d_pad, xsize, ysize, NX, NY, E.wpz_minus_half_dxy, schedule ? d_work : nullptr);
	TW_LAUNCH_CHECK(ctx);
	if (schedule) {
		unsigned const max_work = (unsigned)padded_elems;
		order_hist_kernel<<<(nt + 255)/256, 256, 0, st>>>(d_work, nt, max_work, d_hist);
		TW_LAUNCH_CHECK(ctx);
		order_scan_kernel<<<1, 32, 0, st>>>(d_hist);
		TW_LAUNCH_CHECK(ctx);
		order_scatter_kernel<<<(nt + 255)/256, 256, 0, st>>>(d_work, nt, max_work, d_hist, d_order);
		TW_LAUNCH_CHECK(ctx);
	}
	else {d_order = nullptr;}
	switch (pick_group(nt)) {
	case 1:  launch_droplets<1 >(st, d_pad, nt, xsize, ysize, num_iters, E, ctx->d_dir_table, d_steps, d_order); break;
	case 2:  launch_droplets<2 >(st, d_pad, nt, xsize, ysize, num_iters, E, ctx->d_dir_table, d_steps, d_order); break;
	case 4:  launch_droplets<4 >(st, d_pad, nt, xsize, ysize, num_iters, E, ctx->d_dir_table, d_steps, d_order); break;
	case 8:  launch_droplets<8 >(st, d_pad, nt, xsize, ysize, num_iters, E, ctx->d_dir_table, d_steps, d_order); break;
	case 16: launch_droplets<16>(st, d_pad, nt, xsize, ysize, num_iters, E, ctx->d_dir_table, d_steps, d_order); break;
	default: launch_droplets<32>(st, d_pad, nt, xsize, ysize, num_iters, E, ctx->d_dir_table, d_steps, d_order); break;
	}
	TW_LAUNCH_CHECK(ctx);
	unpad_kernel<<<dim3((xsize + 255)/256, ysize, nt), 256, 0, st>>>(d_pad, maps, xsize, ysize, NX, NY, d_min_zvals, min_zval_all);
	TW_LAUNCH_CHECK(ctx);
	return TW_OK;
}

// largest chunk (heightmaps per enqueue) whose scratch fits in `budget` bytes
uint32_t twi_erode_chunk_for(size_t budget, uint32_t ntiles, int xsize, int ysize) {
	size_t const per = (size_t)(xsize + 2*PAD)*(ysize + 2*PAD)*sizeof(float) + 2*sizeof(unsigned);
	size_t c = budget/per;
	if (c < 1) c = 1;
	if (c > ntiles) c = ntiles;
	if (c > 65535) c = 65535; // gridDim.z limit
	return (uint32_t)c;
}

// tw_erode_parallel: `num_threads` droplets of ONE heightmap in flight (0 = as many groups as keep the GPU busy)
int twi_erode_parallel(tw_ctx *ctx, float *d_map, int xsize, int ysize, float min_zval, uint32_t num_iters, const tw_erosion_params *p, uint32_t num_threads) {
	ctx->last_erosion_steps = 0;
	if (num_iters == 0 || p->erode_amount <= 0.0) return TW_OK; // erosion disabled, src/erosion.cpp:16
	if (xsize <= 0 || ysize <= 0) return tw_set_error(ctx, TW_ERR_ARG, "tw_erode_parallel: empty heightmap");
	if (!ctx->d_dir_table) return tw_set_error(ctx, TW_ERR_STATE, "tw_set_sin_table() has not been called");
	int rc = tw_reserve(ctx, 2, 4096);
	if (rc) return rc;
	unsigned long long *d_steps = (unsigned long long *)((char *)ctx->d_scratch[2] + 2048);
	unsigned *d_next = (unsigned *)(d_steps + 1);
	TW_CUDA(ctx, cudaMemsetAsync(d_steps, 0, 16, ctx->stream));
	int const NX = xsize + 2*PAD, NY = ysize + 2*PAD;
	rc = tw_reserve(ctx, 1, (size_t)NX*NY*sizeof(float));
	if (rc) return rc;
	float *d_pad = (float *)ctx->d_scratch[1];
	EParams const E = make_eparams(p);
	cudaStream_t const st = ctx->stream;
	pad_kernel<<<dim3((NX + 255)/256, NY, 1), 256, 0, st>>>(d_map, d_pad, xsize, ysize, NX, NY, E.wpz_minus_half_dxy, nullptr);
	TW_LAUNCH_CHECK(ctx);
	unsigned groups = num_threads ? num_threads : 65536u; // auto: the 65536-map operating point of pick_group() (8 lanes per droplet)
	if (groups > num_iters) {groups = num_iters;}
	switch (pick_group(groups)) {
	case 1:  launch_droplets_shared<1 >(st, d_pad, groups, xsize, ysize, num_iters, E, ctx->d_dir_table, d_steps, d_next); break;
	case 2:  launch_droplets_shared<2 >(st, d_pad, groups, xsize, ysize, num_iters, E, ctx->d_dir_table, d_steps, d_next); break;
	case 4:  launch_droplets_shared<4 >(st, d_pad, groups, xsize, ysize, num_iters, E, ctx->d_dir_table, d_steps, d_next); break;
	case 8:  launch_droplets_shared<8 >(st, d_pad, groups, xsize, ysize, num_iters, E, ctx->d_dir_table, d_steps, d_next); break;
	case 16: launch_droplets_shared<16>(st, d_pad, groups, xsize, ysize, num_iters, E, ctx->d_dir_table, d_steps, d_next); break;
	default: launch_droplets_shared<32>(st, d_pad, groups, xsize, ysize, num_iters, E, ctx->d_dir_table, d_steps, d_next); break;
	}
	TW_LAUNCH_CHECK(ctx);
	unpad_kernel<<<dim3((xsize + 255)/256, ysize, 1), 256, 0, st>>>(d_pad, d_map, xsize, ysize, NX, NY, nullptr, min_zval);
	TW_LAUNCH_CHECK(ctx);
	unsigned long long h_steps = 0;
	TW_CUDA(ctx, cudaMemcpyAsync(&h_steps, d_steps, sizeof(h_steps), cudaMemcpyDeviceToHost, st));
	TW_CUDA(ctx, cudaStreamSynchronize(st));
	ctx->last_erosion_steps = h_steps;
	return TW_OK;
}

int twi_erode(tw_ctx *ctx, float *d_maps, uint32_t ntiles, int xsize, int ysize, const float *d_min_zvals, float min_zval_all,
              uint32_t num_iters, const tw_erosion_params *p)
{
	ctx->last_erosion_steps = 0;
	if (num_iters == 0 || p->erode_amount <= 0.0) return TW_OK; // erosion disabled, src/erosion.cpp:16
	if (xsize <= 0 || ysize <= 0 || ntiles == 0) return tw_set_error(ctx, TW_ERR_ARG, "tw_erode: empty heightmap");
	if (!ctx->d_dir_table) return tw_set_error(ctx, TW_ERR_STATE, "tw_set_sin_table() has not been called");
	int rc = tw_reserve(ctx, 2, 4096);
	if (rc) return rc;
	unsigned long long *d_steps = (unsigned long long *)((char *)ctx->d_scratch[2] + 2048);
	TW_CUDA(ctx, cudaMemsetAsync(d_steps, 0, sizeof(unsigned long long), ctx->stream));
	// process heightmaps in chunks so that the padded scratch stays within a third of the free device memory (each chunk has its own
	// heaviest-first schedule and its own tail, so fewer, larger chunks are better)
	size_t free_b = 0, total_b = 0;
	TW_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
	size_t budget = (free_b + ctx->scratch_bytes[1])/3;
	if (budget < ((size_t)1 << 30)) budget = (size_t)1 << 30;
	uint32_t const chunk = twi_erode_chunk_for(budget, ntiles, xsize, ysize);
	rc = tw_reserve(ctx, 1, twi_erode_scratch_bytes(chunk, xsize, ysize));
	if (rc) return rc;
	for (uint32_t t0 = 0; t0 < ntiles; t0 += chunk) {
		uint32_t const nt = (ntiles - t0 < chunk) ? (ntiles - t0) : chunk;
		rc = twi_erode_enqueue(ctx, ctx->stream, ctx->d_scratch[1], chunk, d_maps + (size_t)t0*xsize*ysize, nt, xsize, ysize,
		                       d_min_zvals ? d_min_zvals + t0 : nullptr, min_zval_all, num_iters, p, d_steps);
		if (rc) return rc;
	}
	unsigned long long h_steps = 0;
	TW_CUDA(ctx, cudaMemcpyAsync(&h_steps, d_steps, sizeof(h_steps), cudaMemcpyDeviceToHost, ctx->stream));
	TW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->last_erosion_steps = h_steps;
	return TW_OK;
}
