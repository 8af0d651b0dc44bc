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

namespace {

constexpr int PAD = 4;

__device__ __forceinline__ float smin(float a, float b) {return (b < a) ? b : a;} // std::min
__device__ __forceinline__ float smax(float a, float b) {return (a < b) ? b : a;} // std::max
__device__ __forceinline__ int clampi(int v, int hi) {return max(min(v, hi), 0);}

struct EParams {
	float erode_amount, wpz_minus_half_dxy, zmin, zrange, relh_adj_tex, clip_hd1;
};

__global__ void pad_kernel(const float *__restrict__ in, float *__restrict__ out, int xsize, int ysize, int NX, int NY) {
	int const x = blockIdx.x*blockDim.x + threadIdx.x, y = blockIdx.y;
	size_t const tile = blockIdx.z;
	if (x >= NX) return;
	int const sx = clampi(x - PAD, xsize - 1), sy = clampi(y - PAD, ysize - 1);
	out[tile*NX*NY + (size_t)y*NX + x] = __ldg(in + tile*xsize*ysize + (size_t)sy*xsize + sx);
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

__global__ void __launch_bounds__(128)
droplet_kernel(float *__restrict__ padded, unsigned ntiles, int xsize, int ysize, unsigned num_iters, EParams E,
	const float2 *__restrict__ dir_table, unsigned long long *__restrict__ steps_out)
{
	unsigned const warp = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	int const lane = threadIdx.x & 31;
	if (warp >= ntiles) return;
	int const NX = xsize + 2*PAD, NY = ysize + 2*PAD;
	float *mh = padded + (size_t)warp*NX*NY;
	float const Kq=10, Kw=0.001f, Kr=0.9f, Kd=0.02f, Ki=0.1f, minSlope=0.05f, g=20, Kg=g*2;
	unsigned const MAX_PATH_LEN = 4u*(unsigned)NX*(unsigned)NY;
	float const erode_amount = E.erode_amount;
	unsigned long long steps = 0;
	// lane-constant brush / deposit offsets
	int const bx = (lane & 3) - 1, bz = ((lane >> 2) & 3) - 1;  // brush cell offsets for lanes 0..15
	int const cx = lane & 1, cz = (lane >> 1) & 1;               // deposit corner for lanes 0..3

#define HMAP(x, y) mh[(size_t)NX*clampi((y), NY-1) + clampi((x), NX-1)]
	// DEPOSIT(H): src/erosion.cpp:42-54; lanes 0-3 own one corner each (inside cells are distinct => no aliasing)
#define DEPOSIT(H) { \
	if (lane < 4) { \
		int const X = xi + cx, Z = zi + cz; \
		float const W = (cx ? xf : (1-xf))*(cz ? zf : (1-zf)); \
		float const delta = ds*erode_amount*W; \
		if (!(X < 0 || Z < 0 || X >= NX || Z >= NY)) {mh[(size_t)NX*Z + X] += delta;} \
	} \
	__syncwarp(); \
	(H) += ds; }

	for (unsigned iter = 0; iter < num_iters; ++iter) {
		Rng rgen; rgen.s1 = (int)iter + 11; rgen.s2 = 79*(int)iter + 121;
		int xi = PAD + (rgen.rand()%xsize);
		int zi = PAD + (rgen.rand()%ysize);
		float xp=xi, zp=zi, xf=0, zf=0, s=0, v=0, w=1, dx=0, dz=0;
		float h=HMAP(xi, zi), h00=h, h10=HMAP(xi+1, zi), h01=HMAP(xi, zi+1), h11=HMAP(xi+1, zi+1);

		for (unsigned numMoves = 0; numMoves < MAX_PATH_LEN; ++numMoves) {
			++steps;
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
			int const nxi=tw_x86_f2i(floorf(nxp)), nzi=tw_x86_f2i(floorf(nzp)); // x86 semantics: NaN -> INT_MIN -> "outside" next step
			float const nxf=nxp-(float)nxi, nzf=nzp-(float)nzi;
			float const nh00=HMAP(nxi, nzi), nh10=HMAP(nxi+1, nzi), nh01=HMAP(nxi, nzi+1), nh11=HMAP(nxi+1, nzi+1);
			float const nh=(nh00*(1-nxf)+nh10*nxf)*(1-nzf)+(nh01*(1-nxf)+nh11*nxf)*nzf;
			if (smax(smax(nh00, nh10), smax(nh01, nh11)) < E.wpz_minus_half_dxy) break; // reached ocean water

			bool const outside = (xi < 0 || zi < 0 || xi >= NX || zi >= NY);
			if (nh>=h || outside) {
				float ds=(nh-h)+0.001f;
				if (ds>=s || outside) {
					ds=s;
					DEPOSIT(h)
					s=0;
					break;
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
				bool const interior = (xi >= 1 && zi >= 1 && xi + 2 <= NX - 1 && zi + 2 <= NY - 1);
				if (interior) { // 16 distinct cells: one lane each
					if (lane < 16) {
						int const x = xi + bx, z = zi + bz;
						float const zo=(float)z-zp, zo2=zo*zo, xo=(float)x-xp;
						float wgt=1-(xo*xo+zo2)*0.25f;
						if (!(wgt<=0)) {
							wgt*=0.1591549430918953f;
							float const delta=ds*erode_amount*wgt;
							mh[(size_t)NX*z + x]-=delta;
						}
					}
				}
				else if (lane == 0) { // border: clamped indices may alias, keep the reference's serial order
					for (int z=zi-1; z<=zi+2; ++z) {
						float const zo=(float)z-zp, zo2=zo*zo;
						for (int x=xi-1; x<=xi+2; ++x) {
							float const xo=(float)x-xp;
							float wgt=1-(xo*xo+zo2)*0.25f;
							if (wgt<=0) continue;
							wgt*=0.1591549430918953f;
							float const delta=ds*erode_amount*wgt;
							HMAP(x, z)-=delta;
						}
					}
				}
				__syncwarp();
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
#undef DEPOSIT
	if (lane == 0 && steps_out) {atomicAdd(steps_out, steps);}
}

} // namespace

int twi_erode(tw_ctx *ctx, float *d_maps, uint32_t ntiles, int xsize, int ysize, const float *d_min_zvals, float min_zval_all,
              uint32_t num_iters, const tw_erosion_params *p)
{
	ctx->last_erosion_steps = 0;
	if (num_iters == 0 || p->erode_amount <= 0.0) return TW_OK; // erosion disabled, src/erosion.cpp:16
	if (xsize <= 0 || ysize <= 0 || ntiles == 0) return tw_set_error(ctx, TW_ERR_ARG, "tw_erode: empty heightmap");
	if (!ctx->d_dir_table) return tw_set_error(ctx, TW_ERR_STATE, "tw_set_sin_table() has not been called");
	int const NX = xsize + 2*PAD, NY = ysize + 2*PAD;
	size_t const padded_elems = (size_t)NX*NY;
	EParams E;
	E.erode_amount = p->erode_amount;
	E.wpz_minus_half_dxy = p->water_plane_z - p->half_dxy;
	E.zmin = p->zmin; E.zrange = p->zmax - p->zmin;
	E.relh_adj_tex = p->relh_adj_tex; E.clip_hd1 = p->clip_hd1;

	int rc = tw_reserve(ctx, 2, 4096);
	if (rc) return rc;
	unsigned long long *d_steps = (unsigned long long *)((char *)ctx->d_scratch[2] + 2048);
	TW_CUDA(ctx, cudaMemsetAsync(d_steps, 0, sizeof(unsigned long long), ctx->stream));

	// process tiles in chunks so that the padded scratch stays below ~4 GiB
	size_t const max_chunk_bytes = (size_t)4 << 30;
	uint32_t chunk = (uint32_t)(max_chunk_bytes/(padded_elems*sizeof(float)));
	if (chunk < 1) chunk = 1;
	if (chunk > ntiles) chunk = ntiles;
	if (chunk > 65535) chunk = 65535; // gridDim.z limit
	rc = tw_reserve(ctx, 1, (size_t)chunk*padded_elems*sizeof(float));
	if (rc) return rc;
	float *d_pad = (float *)ctx->d_scratch[1];
	for (uint32_t t0 = 0; t0 < ntiles; t0 += chunk) {
		uint32_t const nt = (ntiles - t0 < chunk) ? (ntiles - t0) : chunk;
		float *maps = d_maps + (size_t)t0*xsize*ysize;
		pad_kernel<<<dim3((NX + 255)/256, NY, nt), 256, 0, ctx->stream>>>(maps, d_pad, xsize, ysize, NX, NY);
		TW_LAUNCH_CHECK(ctx);
		unsigned const warps_per_block = 4;
		droplet_kernel<<<(nt + warps_per_block - 1)/warps_per_block, 32*warps_per_block, 0, ctx->stream>>>(d_pad, nt, xsize, ysize, num_iters, E, ctx->d_dir_table, d_steps);
		TW_LAUNCH_CHECK(ctx);
		unpad_kernel<<<dim3((xsize + 255)/256, ysize, nt), 256, 0, ctx->stream>>>(d_pad, maps, xsize, ysize, NX, NY, d_min_zvals ? d_min_zvals + t0 : nullptr, min_zval_all);
		TW_LAUNCH_CHECK(ctx);
	}
	unsigned long long h_steps = 0;
	TW_CUDA(ctx, cudaMemcpyAsync(&h_steps, d_steps, sizeof(h_steps), cudaMemcpyDeviceToHost, ctx->stream));
	TW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->last_erosion_steps = h_steps;
	return TW_OK;
}
