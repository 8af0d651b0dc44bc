/* TEST INFRASTRUCTURE ONLY - see terrain_oracle.h. CPU restatement of the reference terrain hot path, pinned bit-for-bit against
 * the unmodified reference objects (oracle/_ref). Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC.
 * "ref:" comments give the reference file:line (relative to the 3DWorld root) each block follows. */
#include "terrain_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ RNG (ref: src/rand_gen.h:22-26) */
static void rng_step(tw_rng *r) {
	long s1 = (long)r->rseed1, s2 = (long)r->rseed2;
	if ((s1 = 40014*(s1%53668) - 12211*(s1/53668)) < 0) s1 += 2147483563;
	if ((s2 = 40692*(s2%52774) - 3791 *(s2/52774)) < 0) s2 += 2147483399;
	r->rseed1 = s1; r->rseed2 = s2;
}
void to_rng_set(tw_rng *r, long s1, long s2) {r->rseed1 = s1; r->rseed2 = s2;}
int to_rng_rand(tw_rng *r) { /* ref: src/rand_gen.h:66-70, T=int */
	rng_step(r);
	int v = (int)r->rseed1 - (int)r->rseed2;
	if (v < 1) v += 2147483562;
	return v;
}
double to_rng_randd(tw_rng *r) { /* ref: src/gen_object.cpp:377-381, T=double */
	rng_step(r);
	double v = (double)r->rseed1 - (double)r->rseed2;
	if (v < 1) v += 2147483562;
	return v/2147483563.;
}
float to_rng_rand_float(tw_rng *r) {return (float)(0.000001*(to_rng_rand(r)%1000000));} /* ref: src/rand_gen.h:86 */
float to_rng_rand_uniform(tw_rng *r, float a, float b) {return a + (b - a)*(float)to_rng_randd(r);} /* ref: src/rand_gen.h:90 */

/* ------------------------------------------------------------------ sin table (ref: src/sinf.h:8-20, src/mesh_gen.cpp:72-81) */
#define TSIZE 32768
static const float PI_F = 3.141592654f;                 /* ref: src/3DWorld.h:43 */
static float two_pi(void) {return (float)(2.0*PI_F);}   /* ref: src/3DWorld.h:129 */
static float sscale_f(void) {return (float)TSIZE/two_pi();} /* ref: src/sinf.h:9 */

void to_build_sin_table(float *tab) {
	float const sscale = sscale_f();
	for (unsigned i = 0; i < TSIZE; ++i) {
		tab[i]       = sinf((float)i/sscale);
		tab[i+TSIZE] = cosf((float)i/sscale);
	}
}
float to_sinf_lut(const float *tab, float v) {
	float const sscale = sscale_f();
	return (v < 0) ? -tab[((int)(sscale*(-v)))&(TSIZE-1)] : tab[((int)(sscale*v))&(TSIZE-1)];
}
float to_cosf_lut(const float *tab, float v) {
	float const sscale = sscale_f();
	return tab[TSIZE + (((int)(sscale*fabsf(v)))&(TSIZE-1))];
}

/* ------------------------------------------------------------------ host-side parameter generation */
int to_compute_scale(float mesh_scale, int mesh_freq_filter) { /* ref: src/mesh_gen.cpp:544-548 */
	int const iscale = (int)log2f(mesh_scale);
	int v = iscale + mesh_freq_filter;
	if (v > 9-3) v = 9-3;
	if (v < 0) v = 0;
	return 10*v;
}

static void apply_mesh_rand_seed(tw_rng *r, int mesh_seed, int mesh_rgen_index, int mode) { /* ref: src/mesh_gen.cpp:213-216 */
	if (mesh_seed != 0) {to_rng_set(r, mesh_seed, 12345);}
	else if (mode != TW_MGEN_SINE) {to_rng_set(r, mesh_rgen_index+1, 12345);}
}

void to_gen_sine_params(tw_rng *rgen, float scaled_height, int MX, int MY, float XSS, float YSS, int mesh_seed, int mesh_rgen_index,
	int mode, float start_mag, float start_freq, float mag_mult, float freq_mult, float *T) /* ref: src/mesh_gen.cpp:219-254 */
{
	float xf_scale = (float)MY/(float)MX, yf_scale = (float)(1.0/xf_scale);
	if (XSS > YSS) yf_scale *= (float)YSS/(float)XSS;
	if (YSS > XSS) xf_scale *= (float)XSS/(float)YSS;
	float mags[9], freqs[9];
	freqs[0] = start_freq; mags[0] = start_mag;
	for (int i = 1; i < 9; ++i) {freqs[i] = freqs[i-1]*freq_mult; mags[i] = mags[i-1]*mag_mult;}
	float const mesh_h = (float)(scaled_height/sqrt(0.1*10));
	apply_mesh_rand_seed(rgen, mesh_seed, mesh_rgen_index, mode);
	for (int l = 0; l < 9; ++l) {
		float const x_freq = freqs[l]/((float)MX), y_freq = freqs[l]/((float)MY);
		float const mheight = mags[l]*mesh_h;
		for (int i = 0; i < 10; ++i) {
			float *e = T + 5*(l*10 + i);
			e[0] = to_rng_rand_uniform(rgen, 0.2f, 1.0f)*mheight;
			e[1] = to_rng_rand_float(rgen)*two_pi();
			e[2] = to_rng_rand_float(rgen)*two_pi();
			e[3] = to_rng_rand_uniform(rgen, 0.1f, 1.0f)*x_freq*yf_scale;
			e[4] = to_rng_rand_uniform(rgen, 0.1f, 1.0f)*y_freq*xf_scale;
		}
	}
}

void to_gen_rx_ry(int mesh_seed, int mesh_rgen_index, int mode, float *rx, float *ry) { /* ref: src/mesh_gen.cpp:581-586 */
	tw_rng r; to_rng_set(&r, 1, 1);
	apply_mesh_rand_seed(&r, mesh_seed, mesh_rgen_index, mode);
	*rx = (float)(to_rng_rand_float(&r) + 1.0);
	*ry = (float)(to_rng_rand_float(&r) + 1.0);
}

static float do_glaciate_exp(float v, float custom) {return (custom == 0.0f) ? v*v*v : powf(v, custom);} /* ref: src/mesh_gen.cpp:358-360 */

float to_water_z_height(float zmax_est, int glaciate, float custom, float water_h_off, float water_h_off_rel) { /* ref: src/mesh_gen.cpp:362,507-512 */
	float t = 0.42f + water_h_off_rel;                   /* W_PLANE_Z */
	float wpz = (t < 1.0f) ? t : 1.0f; wpz = (0.0f < wpz) ? wpz : 0.0f; /* CLIP_TO_01 = max(0.0f, min(1.0f, x)) */
	if (glaciate) {wpz = do_glaciate_exp(wpz, custom);}
	float const zmax_est2 = (float)(2.0*zmax_est);
	return wpz*zmax_est2 - zmax_est + water_h_off;
}

/* ------------------------------------------------------------------ GLM noise (ref: dependencies/glm/glm/detail/_noise.hpp:15-84) */
static inline float mod289f(float x) {return x - floorf(x*(1.0f/289.0f))*289.0f;}
static inline float permutef(float x) {return mod289f(((x*34.0f) + 1.0f)*x);}
static inline float tinvsqrt(float r) {return 1.79284291400159f - 0.85373472095314f*r;}
static inline float fadef(float t) {return (t*t*t)*(t*(t*6.0f - 15.0f) + 10.0f);}
static inline float glm_modf(float a, float b) {return a - b*floorf(a/b);}     /* ref: detail/func_common.inl:216 */
static inline float fractf(float x) {return x - floorf(x);}                     /* ref: detail/func_common.inl:188 */
static inline float mixf(float x, float y, float a) {return x + a*(y - x);}      /* ref: detail/func_common.inl:129 */
static inline float glm_max(float x, float y) {return (x < y) ? y : x;}          /* ref: detail/func_common.inl:28 */
static inline float glm_min(float x, float y) {return (y < x) ? y : x;}          /* ref: detail/func_common.inl:19 */
static inline float glm_step(float edge, float x) {return (x < edge) ? 0.0f : 1.0f;} /* ref: detail/func_common.inl:252,545 */

float to_simplex2(float vx, float vy) { /* ref: gtc/noise.inl:592-646 */
	float const Cx = 0.211324865405187f, Cy = 0.366025403784439f, Cz = -0.577350269189626f, Cw = 0.024390243902439f;
	float const s = vx*Cy + vy*Cy;
	float ix = floorf(vx + s), iy = floorf(vy + s);
	float const t = ix*Cx + iy*Cx;
	float const x0x = vx - ix + t, x0y = vy - iy + t;
	float const i1x = (x0x > x0y) ? 1.0f : 0.0f, i1y = (x0x > x0y) ? 0.0f : 1.0f;
	float x12x = x0x + Cx, x12y = x0y + Cx, x12z = x0x + Cz, x12w = x0y + Cz;
	x12x = x12x - i1x; x12y = x12y - i1y;
	ix = glm_modf(ix, 289.0f); iy = glm_modf(iy, 289.0f);
	float const q0 = permutef(iy + 0.0f), q1 = permutef(iy + i1y), q2 = permutef(iy + 1.0f);
	float const p0 = permutef(q0 + ix + 0.0f), p1 = permutef(q1 + ix + i1x), p2 = permutef(q2 + ix + 1.0f);
	float m0 = glm_max(0.5f - (x0x*x0x + x0y*x0y), 0.0f);
	float m1 = glm_max(0.5f - (x12x*x12x + x12y*x12y), 0.0f);
	float m2 = glm_max(0.5f - (x12z*x12z + x12w*x12w), 0.0f);
	m0 = m0*m0; m1 = m1*m1; m2 = m2*m2;
	m0 = m0*m0; m1 = m1*m1; m2 = m2*m2;
	float const X0 = 2.0f*fractf(p0*Cw) - 1.0f, X1 = 2.0f*fractf(p1*Cw) - 1.0f, X2 = 2.0f*fractf(p2*Cw) - 1.0f;
	float const h0 = fabsf(X0) - 0.5f, h1 = fabsf(X1) - 0.5f, h2 = fabsf(X2) - 0.5f;
	float const ox0 = floorf(X0 + 0.5f), ox1 = floorf(X1 + 0.5f), ox2 = floorf(X2 + 0.5f);
	float const a00 = X0 - ox0, a01 = X1 - ox1, a02 = X2 - ox2;
	m0 *= 1.79284291400159f - 0.85373472095314f*(a00*a00 + h0*h0);
	m1 *= 1.79284291400159f - 0.85373472095314f*(a01*a01 + h1*h1);
	m2 *= 1.79284291400159f - 0.85373472095314f*(a02*a02 + h2*h2);
	float const gx = a00*x0x + h0*x0y, gy = a01*x12x + h1*x12y, gz = a02*x12z + h2*x12w;
	return 130.0f*(m0*gx + m1*gy + m2*gz);
}

float to_perlin2(float Px, float Py) { /* ref: gtc/noise.inl:25-62 */
	float const flx = floorf(Px), fly = floorf(Py);
	float Pix = flx + 0.0f, Piy = fly + 0.0f, Piz = flx + 1.0f, Piw = fly + 1.0f;
	float const frx = fractf(Px), fry = fractf(Py);
	float const Pfx = frx - 0.0f, Pfy = fry - 0.0f, Pfz = frx - 1.0f, Pfw = fry - 1.0f;
	Pix = glm_modf(Pix, 289.0f); Piy = glm_modf(Piy, 289.0f); Piz = glm_modf(Piz, 289.0f); Piw = glm_modf(Piw, 289.0f);
	float const ix[4] = {Pix, Piz, Pix, Piz}, iy[4] = {Piy, Piy, Piw, Piw};
	float const fx[4] = {Pfx, Pfz, Pfx, Pfz}, fy[4] = {Pfy, Pfy, Pfw, Pfw};
	float gx[4], gy[4];
	for (int k = 0; k < 4; ++k) {
		float const i = permutef(permutef(ix[k]) + iy[k]);
		float g = 2.0f*fractf(i/41.0f) - 1.0f;
		gy[k] = fabsf(g) - 0.5f;
		float const tx = floorf(g + 0.5f);
		gx[k] = g - tx;
	}
	/* g00=(gx0,gy0) g10=(gx1,gy1) g01=(gx2,gy2) g11=(gx3,gy3); norm = (g00,g01,g10,g11) */
	float const n_00 = tinvsqrt(gx[0]*gx[0] + gy[0]*gy[0]), n_01 = tinvsqrt(gx[2]*gx[2] + gy[2]*gy[2]);
	float const n_10 = tinvsqrt(gx[1]*gx[1] + gy[1]*gy[1]), n_11 = tinvsqrt(gx[3]*gx[3] + gy[3]*gy[3]);
	float const g00x = gx[0]*n_00, g00y = gy[0]*n_00, g01x = gx[2]*n_01, g01y = gy[2]*n_01;
	float const g10x = gx[1]*n_10, g10y = gy[1]*n_10, g11x = gx[3]*n_11, g11y = gy[3]*n_11;
	float const n00 = g00x*fx[0] + g00y*fy[0];
	float const n10 = g10x*fx[1] + g10y*fy[1];
	float const n01 = g01x*fx[2] + g01y*fy[2];
	float const n11 = g11x*fx[3] + g11y*fy[3];
	float const fdx = fadef(Pfx), fdy = fadef(Pfy);
	float const nx0 = mixf(n00, n10, fdx), nx1 = mixf(n01, n11, fdx);
	float const nxy = mixf(nx0, nx1, fdy);
	return 2.3f*nxy;
}

float to_perlin3(float Px, float Py, float Pz) { /* ref: gtc/noise.inl:66-133 */
	float P[3] = {Px, Py, Pz}, Pi0[3], Pi1[3], Pf0[3], Pf1[3];
	for (int d = 0; d < 3; ++d) {
		float const fl = floorf(P[d]);
		Pi0[d] = mod289f(fl); Pi1[d] = mod289f(fl + 1.0f);
		Pf0[d] = fractf(P[d]); Pf1[d] = Pf0[d] - 1.0f;
	}
	float const ix[4] = {Pi0[0], Pi1[0], Pi0[0], Pi1[0]}, iy[4] = {Pi0[1], Pi0[1], Pi1[1], Pi1[1]};
	float g0[4][3], g1[4][3]; /* [corner][x,y,z] for z0 and z1 planes; corner order 000,100,010,110 */
	for (int k = 0; k < 4; ++k) {
		float const ixy = permutef(permutef(ix[k]) + iy[k]);
		for (int zz = 0; zz < 2; ++zz) {
			float const ixyz = permutef(ixy + (zz ? Pi1[2] : Pi0[2]));
			float gx = ixyz*(float)(1.0/7.0);
			float gy = fractf(floorf(gx)*(float)(1.0/7.0)) - 0.5f;
			gx = fractf(gx);
			float const gz = 0.5f - fabsf(gx) - fabsf(gy);
			float const sz = glm_step(gz, 0.0f);
			gx -= sz*(glm_step(0.0f, gx) - 0.5f);
			gy -= sz*(glm_step(0.0f, gy) - 0.5f);
			float *g = zz ? g1[k] : g0[k];
			g[0] = gx; g[1] = gy; g[2] = gz;
		}
	}
	/* norm0 = taylorInvSqrt(dot(g000),dot(g010),dot(g100),dot(g110)) - each corner scaled by its own norm */
	for (int k = 0; k < 4; ++k) {
		float *g = g0[k]; float n = tinvsqrt(g[0]*g[0] + g[1]*g[1] + g[2]*g[2]); g[0] *= n; g[1] *= n; g[2] *= n;
		g = g1[k];        n = tinvsqrt(g[0]*g[0] + g[1]*g[1] + g[2]*g[2]);       g[0] *= n; g[1] *= n; g[2] *= n;
	}
	float const n000 = g0[0][0]*Pf0[0] + g0[0][1]*Pf0[1] + g0[0][2]*Pf0[2];
	float const n100 = g0[1][0]*Pf1[0] + g0[1][1]*Pf0[1] + g0[1][2]*Pf0[2];
	float const n010 = g0[2][0]*Pf0[0] + g0[2][1]*Pf1[1] + g0[2][2]*Pf0[2];
	float const n110 = g0[3][0]*Pf1[0] + g0[3][1]*Pf1[1] + g0[3][2]*Pf0[2];
	float const n001 = g1[0][0]*Pf0[0] + g1[0][1]*Pf0[1] + g1[0][2]*Pf1[2];
	float const n101 = g1[1][0]*Pf1[0] + g1[1][1]*Pf0[1] + g1[1][2]*Pf1[2];
	float const n011 = g1[2][0]*Pf0[0] + g1[2][1]*Pf1[1] + g1[2][2]*Pf1[2];
	float const n111 = g1[3][0]*Pf1[0] + g1[3][1]*Pf1[1] + g1[3][2]*Pf1[2];
	float const fx = fadef(Pf0[0]), fy = fadef(Pf0[1]), fz = fadef(Pf0[2]);
	float const nz0 = mixf(n000, n001, fz), nz1 = mixf(n100, n101, fz), nz2 = mixf(n010, n011, fz), nz3 = mixf(n110, n111, fz);
	float const nyz0 = mixf(nz0, nz2, fy), nyz1 = mixf(nz1, nz3, fy);
	return 2.2f*mixf(nyz0, nyz1, fx);
}

float to_simplex3(float vx, float vy, float vz) { /* ref: gtc/noise.inl:649-721 */
	float const Cx = (float)(1.0/6.0), Cy = (float)(1.0/3.0);
	float const Dx = 0.0f, Dy = 0.5f, Dz = 1.0f, Dw = 2.0f;
	float const s = vx*Cy + vy*Cy + vz*Cy;
	float i[3] = {floorf(vx + s), floorf(vy + s), floorf(vz + s)};
	float const t = i[0]*Cx + i[1]*Cx + i[2]*Cx;
	float const x0[3] = {vx - i[0] + t, vy - i[1] + t, vz - i[2] + t};
	float const g[3] = {glm_step(x0[1], x0[0]), glm_step(x0[2], x0[1]), glm_step(x0[0], x0[2])};
	float const l[3] = {1.0f - g[0], 1.0f - g[1], 1.0f - g[2]};
	float const lz[3] = {l[2], l[0], l[1]};
	float i1[3], i2[3], x1[3], x2[3], x3[3];
	for (int d = 0; d < 3; ++d) {
		i1[d] = glm_min(g[d], lz[d]); i2[d] = glm_max(g[d], lz[d]);
		x1[d] = x0[d] - i1[d] + Cx; x2[d] = x0[d] - i2[d] + Cy; x3[d] = x0[d] - Dy;
	}
	for (int d = 0; d < 3; ++d) {i[d] = mod289f(i[d]);}
	float const oz[4] = {0.0f, i1[2], i2[2], 1.0f}, oy[4] = {0.0f, i1[1], i2[1], 1.0f}, ox[4] = {0.0f, i1[0], i2[0], 1.0f};
	float p[4];
	for (int k = 0; k < 4; ++k) {p[k] = permutef(permutef(permutef(i[2] + oz[k]) + i[1] + oy[k]) + i[0] + ox[k]);}
	float const n_ = 0.142857142857f;
	float const nsx = n_*Dw - Dx, nsy = n_*Dy - Dz, nsz = n_*Dz - Dx;
	float X[4], Y[4], H[4];
	for (int k = 0; k < 4; ++k) {
		float const j  = p[k] - 49.0f*floorf(p[k]*nsz*nsz);
		float const x_ = floorf(j*nsz);
		float const y_ = floorf(j - 7.0f*x_);
		X[k] = x_*nsx + nsy; Y[k] = y_*nsx + nsy;
		H[k] = 1.0f - fabsf(X[k]) - fabsf(Y[k]);
	}
	float const b0[4] = {X[0], X[1], Y[0], Y[1]}, b1[4] = {X[2], X[3], Y[2], Y[3]};
	float s0[4], s1[4], sh[4];
	for (int k = 0; k < 4; ++k) {s0[k] = floorf(b0[k])*2.0f + 1.0f; s1[k] = floorf(b1[k])*2.0f + 1.0f; sh[k] = -glm_step(H[k], 0.0f);}
	float const a0[4] = {b0[0] + s0[0]*sh[0], b0[2] + s0[2]*sh[0], b0[1] + s0[1]*sh[1], b0[3] + s0[3]*sh[1]};
	float const a1[4] = {b1[0] + s1[0]*sh[2], b1[2] + s1[2]*sh[2], b1[1] + s1[1]*sh[3], b1[3] + s1[3]*sh[3]};
	float P0[3] = {a0[0], a0[1], H[0]}, P1[3] = {a0[2], a0[3], H[1]}, P2[3] = {a1[0], a1[1], H[2]}, P3[3] = {a1[2], a1[3], H[3]};
	float *PP[4] = {P0, P1, P2, P3};
	for (int k = 0; k < 4; ++k) {
		float *q = PP[k]; float const n = tinvsqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2]);
		q[0] *= n; q[1] *= n; q[2] *= n;
	}
	float const *XX[4] = {x0, x1, x2, x3};
	float m[4], d[4];
	for (int k = 0; k < 4; ++k) {
		float const *x = XX[k];
		m[k] = glm_max(0.6f - (x[0]*x[0] + x[1]*x[1] + x[2]*x[2]), 0.0f);
		m[k] = m[k]*m[k];
		d[k] = PP[k][0]*x[0] + PP[k][1]*x[1] + PP[k][2]*x[2];
	}
	float const mm0 = m[0]*m[0], mm1 = m[1]*m[1], mm2 = m[2]*m[2], mm3 = m[3]*m[3];
	return 42.0f*((mm0*d[0] + mm1*d[1]) + (mm2*d[2] + mm3*d[3]));
}

/* ------------------------------------------------------------------ 2-D height path */
static inline float std_min(float a, float b) {return (b < a) ? b : a;} /* std::min */
static inline float std_max(float a, float b) {return (a < b) ? b : a;} /* std::max */

static void postproc_noise_zval(float *zval, const tw_hmap_params *h) { /* ref: src/mesh_gen.cpp:555-562 */
	float z = *zval;
	if (z > h->plat_bot) {z = h->plat_bot + h->plat_h*(z - h->plat_bot) + std_min(h->plat_max, h->plat_s*(z - h->plat_bot));}
	if (z > h->crat_h  ) {z = h->crat_h - h->crat_s*(z - h->crat_h);}
	if (z > h->crack_lo && z < h->crack_hi) {z -= h->crack_d*std_min(z - h->crack_lo, h->crack_hi - z);}
	*zval = z;
}
static void apply_noise_shape_final(float *noise, int shape, const tw_hmap_params *h) { /* ref: src/mesh_gen.cpp:564-571 */
	switch (shape) {
	case 1: *noise = (float)(fabsf(*noise) - 2.0); break;
	case 2: *noise = (float)(3.5 - fabsf(*noise)); break;
	default: break;
	}
	postproc_noise_zval(noise, h);
}
static float get_hmap_scale(const tw_height_params *p) { /* ref: src/mesh_gen.cpp:550-553 */
	int const m = p->gen_mode;
	float const scale = (m == TW_MGEN_SIMPLEX || m == TW_MGEN_SIMPLEX_GPU || m == TW_MGEN_DWARP_GPU) ? 16.0f : 32.0f;
	return scale*p->mesh_height*p->mesh_height_scale*p->mesh_scale_z_inv;
}
static float gen_noise(float xv, float yv, const tw_height_params *p) { /* ref: src/mesh_gen.cpp:706-730 */
	float zval = 0.0f, mag = 1.0f, freq = 1.0f, rx = p->rx, ry = p->ry;
	unsigned const end_octave = 9 - p->start_eval_sin/10;
	float const lacunarity = 1.92f, gain = 0.5f;
	int const mode = p->gen_mode;
	for (unsigned i = 0; i < end_octave; ++i) {
		float const px = freq*xv + rx, py = freq*yv + ry;
		float noise = (mode == TW_MGEN_SIMPLEX || mode == TW_MGEN_SIMPLEX_GPU || mode == TW_MGEN_DWARP_GPU) ? to_simplex2(px, py) : to_perlin2(px, py);
		switch (p->gen_shape) {
		case 1: noise = (float)(fabsf(noise) - 0.40); break;
		case 2: noise = (float)(0.45 - fabsf(noise)); break;
		default: break;
		}
		zval += mag*noise;
		mag  *= gain;
		freq *= lacunarity;
		rx   *= 1.5f;
		ry   *= 1.5f;
	}
	return zval;
}
float to_get_noise_zval(float xval, float yval, const tw_height_params *p) { /* ref: src/mesh_gen.cpp:734-751 */
	float const xy_scale = 0.0007f*p->mesh_scale;
	float xv = xy_scale*xval, yv = xy_scale*yval;
	if (p->gen_mode == TW_MGEN_DWARP_GPU) {
		float const scale = 0.2f;
		float const dx1 = gen_noise((float)(xv + 0.0), (float)(yv + 0.0), p);
		float const dy1 = gen_noise((float)(xv + 5.2), (float)(yv + 1.3), p);
		float const dx2 = gen_noise((float)((xv + scale*dx1) + 1.7), (float)((yv + scale*dy1) + 9.2), p);
		float const dy2 = gen_noise((float)((xv + scale*dx1) + 8.3), (float)((yv + scale*dy1) + 2.8), p);
		xv += scale*dx2; yv += scale*dy2;
	}
	float zval = gen_noise(xv, yv, p);
	postproc_noise_zval(&zval, &p->hmap);
	return zval*get_hmap_scale(p);
}
float to_eval_mesh_sin_terms(float xv, float yv, const float *tab, const float *T, int start) { /* ref: src/mesh_gen.cpp:797-805 */
	float zval = 0.0f;
	for (int k = start; k < 90; ++k) {
		float const *s = T + 5*k;
		zval += s[0]*to_sinf_lut(tab, s[3]*yv + s[1])*to_sinf_lut(tab, s[4]*xv + s[2]);
	}
	return zval;
}
static float get_volcano_height(float xi, float yi, const tw_height_params *p, const float *tab) { /* ref: src/mesh_gen.cpp:364-372 */
	float const freq = p->mesh_scale/p->hmap.volcano_width, x = freq*xi, y = freq*yi, dist = sqrtf(x*x + y*y);
	if (dist > 2.0) return 0.0f;
	float const val = to_cosf_lut(tab, x)*to_cosf_lut(tab, y);
	double const hd = 400.0*(val - 0.999);
	float const hole = (float)((0.0 < hd) ? hd : 0.0);
	float const peak = (float)(0.08*val/std_max(0.04f, dist));
	return p->hmap.volcano_height*std_max(0.0f, (peak - hole))*p->mesh_scale_z_inv;
}

void to_heightgen_2d(const tw_grid2d *g, const tw_height_params *p, const float *tab, const float *T, int enable_glaciate,
	int min_start_sin, float *out, int nthreads)
{
	unsigned const nx = g->nx, ny = g->ny;
	float const dx = g->dx, dy = g->dy;
	float const mx0 = dx*g->x0, my0 = dy*g->y0, mdx = dx, mdy = dy; /* ref: src/mesh_gen.cpp:591 */
	int const F = 90, start = p->start_eval_sin;
	float *xy = NULL, *smt = NULL;
	float sine_offset = 0.0f;
	if (p->gen_mode == TW_MGEN_SINE) { /* ref: src/mesh_gen.cpp:604-626 */
		xy = (float *)calloc((size_t)(nx + ny)*F, sizeof(float));
		float const msx = p->mesh_scale*p->dx_val_inv, msy = p->mesh_scale*p->dy_val_inv, ms2 = (float)(0.5*p->mesh_scale);
		for (int k = start; k < F; ++k) {
			float const *s = T + 5*k;
			float const x_mult = msx*s[4], y_mult = msy*s[3], y_scale = p->mesh_scale_z_inv*s[0];
			float const x_const = ms2*s[4] + s[2] + x_mult*mx0, y_const = ms2*s[3] + s[1] + y_mult*my0;
			float const xmdx = x_mult*dx, ymdy = y_mult*dy;
			for (unsigned i = 0; i < nx; ++i) {xy[(size_t)i*F + k] = to_sinf_lut(tab, xmdx*(float)i + x_const);}
			for (unsigned i = 0; i < ny; ++i) {xy[((size_t)nx + i)*F + k] = y_scale*to_sinf_lut(tab, ymdy*(float)i + y_const);}
		}
	}
	if (enable_glaciate && p->hmap.sine_mag != 0.0f) { /* ref: src/mesh_gen.cpp:640-650 */
		smt = (float *)malloc((size_t)(nx + ny)*sizeof(float));
		sine_offset = p->hmap.sine_bias*p->mesh_scale_z_inv;
		float const sm_scale = p->hmap.sine_mag*p->mesh_scale_z_inv, freq = p->mesh_scale*p->hmap.sine_freq;
		for (unsigned x = 0; x < nx; ++x) {smt[x] = sm_scale*to_cosf_lut(tab, ((float)x*mdx + mx0)*p->dx_val_inv*freq);}
		for (unsigned y = 0; y < ny; ++y) {smt[nx + y] = to_cosf_lut(tab, ((float)y*mdy + my0)*p->dy_val_inv*freq);}
	}
	float const zmax_est = p->zmax_est, zmax_est2 = (float)(2.0*zmax_est), zmax_est2_inv = (float)(1.0/zmax_est2); /* ref: :162-167 */
	int const start_ix = (start > min_start_sin) ? start : min_start_sin;
#ifdef _OPENMP
	if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
	(void)nthreads;
#pragma omp parallel for schedule(static,1) num_threads(nthreads)
	for (int y = 0; y < (int)ny; ++y) {
		for (unsigned x = 0; x < nx; ++x) { /* ref: eval_index, src/mesh_gen.cpp:754-792 */
			float zval = 0.0f;
			if (p->gen_mode != TW_MGEN_SINE) {
				float const xval = ((float)x*mdx + mx0)*p->dx_val_inv, yval = ((float)y*mdy + my0)*p->dy_val_inv;
				zval += to_get_noise_zval(xval, yval, p);
			}
			else {
				float const *xptr = xy + (size_t)x*F, *yptr = xy + ((size_t)nx + y)*F;
				for (int i = start_ix; i < F; ++i) {zval += xptr[i]*yptr[i];}
				apply_noise_shape_final(&zval, p->gen_shape, &p->hmap);
			}
			if (enable_glaciate) {
				if (p->glaciate) { /* apply_glaciate, ref: src/mesh_gen.cpp:380-385 */
					float const relh = (zval + zmax_est)*zmax_est2_inv;
					zval = do_glaciate_exp(relh, p->custom_glaciate_exp)*zmax_est2 - zmax_est;
				}
				if (p->hmap.sine_mag > 0.0f) {
					zval += smt[x]*smt[nx + y] + sine_offset;
					if (p->hmap.volcano_width > 0.0f && p->hmap.volcano_height > 0.0f) {
						zval += get_volcano_height(((float)x*mdx + mx0)*p->dx_val_inv, ((float)y*mdy + my0)*p->dy_val_inv, p, tab);
					}
				}
			}
			out[(size_t)y*nx + x] = zval;
		}
	}
	free(xy); free(smt);
}

/* ------------------------------------------------------------------ ground-mode mesh (ref: src/mesh_gen.cpp:257-355,373-404,447-504) */
static void apply_mesh_sine(float *zval, float x, float y, const tw_height_params *p, const float *tab) { /* ref: src/mesh_gen.cpp:373-379 */
	if (p->hmap.sine_mag > 0.0) {
		float const freq = p->mesh_scale*p->hmap.sine_freq;
		*zval += (p->hmap.sine_mag*to_cosf_lut(tab, x*freq)*to_cosf_lut(tab, y*freq) + p->hmap.sine_bias)*p->mesh_scale_z_inv;
		if (p->hmap.volcano_width > 0.0 && p->hmap.volcano_height > 0.0) {*zval += get_volcano_height(x, y, p, tab);}
	}
}
void to_glaciate_mesh(float *mesh, int nx, int ny, int xoff2, int yoff2, int MX, int MY, const tw_height_params *p, const float *tab, float *zbottom, float *ztop)
{
	float const zmax_est = p->zmax_est, zmax_est2 = (float)(2.0*zmax_est), zmax_est2_inv = (float)(1.0/zmax_est2);
	float zb = *zbottom, zt = *ztop;
	for (int i = 0; i < ny; ++i) {
		for (int j = 0; j < nx; ++j) {
			float zval = mesh[(size_t)i*nx + j];
			if (p->glaciate) {
				float const relh = (zval + zmax_est)*zmax_est2_inv;
				zval = do_glaciate_exp(relh, p->custom_glaciate_exp)*zmax_est2 - zmax_est;
			}
			apply_mesh_sine(&zval, (float)(j + xoff2 - MX/2), (float)(i + yoff2 - MY/2), p, tab);
			mesh[(size_t)i*nx + j] = zval;
			zb = std_min(zb, zval);
			zt = std_max(zt, zval);
		}
	}
	*zbottom = zb; *ztop = zt;
}
void to_gen_mesh(tw_rng *sine_rng, tw_height_params *p, int MX, int MY, float XSS, float YSS, int mesh_seed, int mesh_rgen_index, int xoff2, int yoff2,
	float dx_val, float dy_val, float water_h_off, float water_h_off_rel, unsigned erosion_iters, tw_erosion_params *ep, const float *tab,
	float *T, float *mesh, float *zvals6)
{
	float const scaled_height = p->mesh_height*p->mesh_height_scale;              /* :267 */
	float const LARGE_ZVAL = 100.0f*(1.5f*(p->mesh_height/0.10f));               /* only used as the min/max seed */
	to_gen_sine_params(sine_rng, scaled_height, MX, MY, XSS, YSS, mesh_seed, mesh_rgen_index, p->gen_mode, 0.02f, 240.0f, 2.0f, 0.5f, T);
	tw_grid2d g = {(float)(xoff2 - MX/2), (float)(yoff2 - MY/2), dx_val, dy_val, (uint32_t)MX, (uint32_t)MY}; /* gen_mesh_sine_table, :201-210 */
	to_heightgen_2d(&g, p, tab, T, 0, 0, mesh, 1);
	float zmin = mesh[0], zmax = mesh[0];                                             /* calc_zminmax */
	for (size_t i = 0; i < (size_t)MX*MY; ++i) {zmin = std_min(zmin, mesh[i]); zmax = std_max(zmax, mesh[i]);}
	/* estimate_zminmax(using_eq=1), :447-485 */
	float zmax_est = std_max(zmax, -zmin);
	float zbottom, ztop;
	if (zmax == zmin) {zmax_est = (float)(zmax_est + 1.0E-6);}
	else {
		float const XY_SCENE_SIZE = 0.5f*(XSS + YSS);
		float const rm_scale = (float)(1000.0*XY_SCENE_SIZE/p->mesh_scale);
		tw_grid2d ge = {0.0f, 0.0f, rm_scale, rm_scale, 128, 128};
		float *probe = (float *)malloc(128*128*sizeof(float));
		to_heightgen_2d(&ge, p, tab, T, 0, 0, probe, 1);
		for (unsigned i = 0; i < 128*128; ++i) {zmax_est = std_max(zmax_est, (float)fabs(probe[i]));}
		free(probe);
		if (p->gen_mode != TW_MGEN_SINE) {zmax_est = (float)(zmax_est*1.2);}
		zmax_est = (float)(1.1*zmax_est);
	}
	/* set_zvals, :494-504 (the early-return branch above skips it in the reference; zbottom/ztop then keep their previous values - we use zmin/zmax) */
	zbottom = zmin; ztop = zmax;
	zmin = -zmax_est; zmax = zmax_est;
	p->zmax_est = zmax_est;
	float const water_plane_z = to_water_z_height(zmax_est, p->glaciate, p->custom_glaciate_exp, water_h_off, water_h_off_rel);
	/* gen_terrain_map, :434-444 */
	if (p->glaciate) {zbottom = LARGE_ZVAL; ztop = -LARGE_ZVAL; to_glaciate_mesh(mesh, MX, MY, xoff2, yoff2, MX, MY, p, tab, &zbottom, &ztop);}
	ep->water_plane_z = water_plane_z; ep->zmin = zmin; ep->zmax = zmax;
	to_apply_erosion(mesh, MX, MY, zbottom, erosion_iters, ep);
	zvals6[0] = zmin; zvals6[1] = zmax; zvals6[2] = zmax_est; zvals6[3] = zbottom; zvals6[4] = ztop; zvals6[5] = water_plane_z;
}

/* ------------------------------------------------------------------ tile bounds (ref: src/tiled_mesh.cpp:517-540) */
void to_tile_bounds(const float *zvals_all, unsigned ntiles, unsigned zvsize, float wpz_max, float dx_val, float dy_val, unsigned size, tw_tile_bounds *out)
{
	float const FAR_DISTANCE = 100.0f;
	unsigned const block_size = zvsize/4;
	for (unsigned t = 0; t < ntiles; ++t) {
		const float *zvals = zvals_all + (size_t)t*zvsize*zvsize;
		tw_tile_bounds *b = out + t;
		b->mzmin = FAR_DISTANCE; b->mzmax = -FAR_DISTANCE; b->mesh_dz = 0.0f;
		b->wx1 = b->wy1 = 2147483647; b->wx2 = b->wy2 = -1;
		for (unsigned yy = 0; yy < 4; ++yy) {
			for (unsigned xx = 0; xx < 4; ++xx) {
				unsigned const x_end = (xx+1)*block_size, y_end = (yy+1)*block_size;
				float szmin = FAR_DISTANCE, szmax = -FAR_DISTANCE;
				for (unsigned y = yy*block_size; y <= y_end; ++y) {
					for (unsigned x = xx*block_size; x <= x_end; ++x) {
						float const z = zvals[y*zvsize + x];
						szmin = std_min(szmin, z); szmax = std_max(szmax, z);
						if (z < wpz_max) {
							if ((int)x < b->wx1) {b->wx1 = (int)x;}
							if ((int)y < b->wy1) {b->wy1 = (int)y;}
							if ((int)x > b->wx2) {b->wx2 = (int)x;}
							if ((int)y > b->wy2) {b->wy2 = (int)y;}
						}
					}
				}
				b->sub_zmin[yy*4 + xx] = szmin; b->sub_zmax[yy*4 + xx] = szmax;
				b->mesh_dz = std_max(b->mesh_dz, (szmax - szmin)); /* max_eq */
				b->mzmin = std_min(b->mzmin, szmin);
				b->mzmax = std_max(b->mzmax, szmax);
			}
		}
		b->radius = (float)(0.5*sqrtf((dx_val*dx_val + dy_val*dy_val)*size*size + (b->mzmax - b->mzmin)*(b->mzmax - b->mzmin)));
	}
}

/* ------------------------------------------------------------------ erosion (ref: src/erosion.cpp:14-164) */
unsigned long long to_apply_erosion(float *heightmap, int xsize, int ysize, float min_zval, unsigned num_iters, const tw_erosion_params *ep)
{
	float const erode_amount = ep->erode_amount;
	if (num_iters == 0 || erode_amount <= 0.0) return 0;
	float const Kq=10, Kw=0.001f, Kr=0.9f, Kd=0.02f, Ki=0.1f, minSlope=0.05f, g=20, Kg=g*2;
	int const PAD = 4, NX = xsize+2*PAD, NY = ysize+2*PAD;
	unsigned const MAX_PATH_LEN = 4*NX*NY;
	float *mh = (float *)malloc((size_t)NX*NY*sizeof(float));
	unsigned long long steps = 0;
	for (int y = 0; y < NY; ++y) {
		int yy = y-PAD; if (yy > ysize-1) yy = ysize-1; if (yy < 0) yy = 0;
		for (int x = 0; x < NX; ++x) {
			int xx = x-PAD; if (xx > xsize-1) xx = xsize-1; if (xx < 0) xx = 0;
			mh[(size_t)y*NX + x] = heightmap[xx + (size_t)yy*xsize];
		}
	}
#define CLAMPI(v, hi) (((v) < (hi)) ? (((v) > 0) ? (v) : 0) : (hi))
#define HMAP_INDEX(x, y) ((size_t)NX*CLAMPI((y), NY-1) + CLAMPI((x), NX-1))
#define HMAP(x, y) mh[HMAP_INDEX(x, y)]
#define DEPOSIT_AT(X, Z, W) { \
	float const delta = ds*erode_amount*(W); \
	size_t const ix = HMAP_INDEX((X), (Z)); \
	if (!((X) < 0 || (Z) < 0 || (X) >= NX || (Z) >= NY)) {mh[ix] += delta;} \
}
#define DEPOSIT(H) \
	DEPOSIT_AT(xi  , zi  , (1-xf)*(1-zf)) \
	DEPOSIT_AT(xi+1, zi  ,    xf *(1-zf)) \
	DEPOSIT_AT(xi  , zi+1, (1-xf)*   zf ) \
	DEPOSIT_AT(xi+1, zi+1,    xf *   zf ) \
	(H)+=ds;
	float const tp = two_pi();
	for (int iter = 0; iter < (int)num_iters; ++iter) {
		tw_rng rgen; to_rng_set(&rgen, iter+11, 79*iter+121);
		int xi = PAD + (to_rng_rand(&rgen)%xsize);
		int zi = PAD + (to_rng_rand(&rgen)%ysize);
		float xp=xi, zp=zi, xf=0, zf=0, s=0, v=0, w=1, dx=0, dz=0;
		float h=HMAP(xi, zi), h00=h, h10=HMAP(xi+1, zi), h01=HMAP(xi, zi+1), h11=HMAP(xi+1, zi+1);
		unsigned numMoves = 0;
		for (; numMoves < MAX_PATH_LEN; ++numMoves) {
			++steps;
			float gx=h00+h01-h10-h11, gz=h00+h10-h01-h11;
			dx=(dx-gx)*Ki+gx;
			dz=(dz-gz)*Ki+gz;
			float dl=sqrtf(dx*dx+dz*dz);
			if (dl<=FLT_EPSILON) {
				float a=to_rng_rand_float(&rgen)*tp;
				dx=cosf(a); dz=sinf(a);
			}
			else {dx/=dl; dz/=dl;}
			float nxp=xp+dx, nzp=zp+dz;
			int nxi=(int)floorf(nxp), nzi=(int)floorf(nzp);
			float nxf=nxp-nxi, nzf=nzp-nzi;
			float nh00=HMAP(nxi, nzi), nh10=HMAP(nxi+1, nzi), nh01=HMAP(nxi, nzi+1), nh11=HMAP(nxi+1, nzi+1);
			float nh=(nh00*(1-nxf)+nh10*nxf)*(1-nzf)+(nh01*(1-nxf)+nh11*nxf)*nzf;
			if (std_max(std_max(nh00, nh10), std_max(nh01, nh11)) < ep->water_plane_z - ep->half_dxy) break;
			int const outside = (xi < 0 || zi < 0 || xi >= NX || zi >= NY);
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
			float slope=dh;
			float q=std_max(slope, minSlope)*v*w*Kq;
			float ds=s-q;
			if (ds>=0) {
				ds*=Kd;
				DEPOSIT(dh)
				s-=ds;
			}
			else {
				ds*=-Kr;
				ds=std_min(ds, dh*0.99f);
				{ /* get_bare_ls_tid(nh) == ROCK_TEX ? 0.5 : 2.0, ref: src/Textures.cpp:1284-1287 */
					float const relh = ep->relh_adj_tex + (nh - ep->zmin)/(ep->zmax - ep->zmin);
					ds = (float)(ds*((relh > ep->clip_hd1) ? 0.5 : 2.0));
				}
				for (int z=zi-1; z<=zi+2; ++z) {
					float zo=z-zp, zo2=zo*zo;
					for (int x=xi-1; x<=xi+2; ++x) {
						float xo=x-xp;
						float wgt=1-(xo*xo+zo2)*0.25f;
						if (wgt<=0) continue;
						wgt*=0.1591549430918953f;
						float const delta=ds*erode_amount*wgt;
						mh[HMAP_INDEX(x, z)]-=delta;
					}
				}
				dh-=ds;
				s+=ds;
			}
			v=sqrtf(v*v+Kg*dh);
			w*=1-Kw;
			xp=nxp; zp=nzp; xi=nxi; zi=nzi; xf=nxf; zf=nzf;
			h=nh; h00=nh00; h10=nh10; h01=nh01; h11=nh11;
		}
	}
	for (int y = 0; y < ysize; ++y) {
		for (int x = 0; x < xsize; ++x) {heightmap[(size_t)y*xsize + x] = std_max(min_zval, mh[(size_t)(y+PAD)*NX + x+PAD]);}
	}
	free(mh);
	return steps;
#undef DEPOSIT_AT
#undef DEPOSIT
#undef HMAP
}

/* ------------------------------------------------------------------ coherent batched erosion (tw_erode_sweeps; NO reference counterpart)
 * The per-move arithmetic is apply_erosion's (above). What differs is what a droplet can see, so that the result does not depend on how the map
 * is split over devices: droplets [k*sweep, (k+1)*sweep) read the map as it was when sweep k began; their deposits go to a 64-bit fixed-point
 * delta buffer (2^-40 units: integer sums are order-independent) that is added to the map after the sweep; a droplet does see its OWN writes
 * through a private VIEW x VIEW window (sweep-start heights + its writes), re-read and re-centred ahead of its heading when it walks out of it;
 * and it ends once it is more than halo - VIEW - 4 rows from its start row or its next position is not finite. Mirrors droplet_kernel<M_FROZEN> in csrc/tw_erosion.cu. */
#define SWEEP_VIEW 32
typedef struct {float *mh; long long *dfix; int NX, NY; float win[SWEEP_VIEW*SWEEP_VIEW]; int WX, WY, wx0, wz0, have;} sweep_state;
static float sw_read(const sweep_state *S, int x, int z) {
	int const cx = CLAMPI(x, S->NX-1), cz = CLAMPI(z, S->NY-1);
	if (S->have) {unsigned const rx = (unsigned)(cx - S->wx0), rz = (unsigned)(cz - S->wz0); if (rx < (unsigned)S->WX && rz < (unsigned)S->WY) return S->win[rz*S->WX + rx];}
	return S->mh[(size_t)S->NX*cz + cx];
}
static void sw_add(sweep_state *S, int x, int z, float delta) { /* (x, z) inside the array */
	if (S->have) {unsigned const rx = (unsigned)(x - S->wx0), rz = (unsigned)(z - S->wz0); if (rx < (unsigned)S->WX && rz < (unsigned)S->WY) {S->win[rz*S->WX + rx] += delta;}}
	if (fabsf(delta) < 1048576.0f) {S->dfix[(size_t)S->NX*z + x] += llrint((double)delta*1099511627776.0);}
}
unsigned long long to_erode_sweeps(float *heightmap, int xsize, int ysize, float min_zval, unsigned num_iters, const tw_erosion_params *ep, unsigned sweep, int halo)
{
	float const erode_amount = ep->erode_amount;
	if (num_iters == 0 || erode_amount <= 0.0 || sweep == 0 || halo < SWEEP_VIEW + 12) return 0;
	float const Kq=10, Kw=0.001f, Kr=0.9f, Kd=0.02f, Ki=0.1f, minSlope=0.05f, g=20, Kg=g*2;
	int const PAD = 4, NX = xsize+2*PAD, NY = ysize+2*PAD, halo_rule = halo - SWEEP_VIEW - PAD;
	unsigned const MAX_PATH_LEN = 4*NX*NY;
	sweep_state S;
	S.mh = (float *)malloc((size_t)NX*NY*sizeof(float));
	S.dfix = (long long *)calloc((size_t)NX*NY, sizeof(long long));
	S.NX = NX; S.NY = NY; S.WX = (SWEEP_VIEW < NX) ? SWEEP_VIEW : NX; S.WY = (SWEEP_VIEW < NY) ? SWEEP_VIEW : NY; S.have = 0; S.wx0 = S.wz0 = 0;
	unsigned long long steps = 0;
	for (int y = 0; y < NY; ++y) {
		int const yy = CLAMPI(y-PAD, ysize-1);
		for (int x = 0; x < NX; ++x) {S.mh[(size_t)y*NX + x] = heightmap[CLAMPI(x-PAD, xsize-1) + (size_t)yy*xsize];}
	}
#define SW_DEPOSIT_AT(X, Z, W) {float const delta = ds*erode_amount*(W); if (!((X) < 0 || (Z) < 0 || (X) >= NX || (Z) >= NY)) {sw_add(&S, (X), (Z), delta);}}
#define SW_DEPOSIT(H) \
	SW_DEPOSIT_AT(xi  , zi  , (1-xf)*(1-zf)) \
	SW_DEPOSIT_AT(xi+1, zi  ,    xf *(1-zf)) \
	SW_DEPOSIT_AT(xi  , zi+1, (1-xf)*   zf ) \
	SW_DEPOSIT_AT(xi+1, zi+1,    xf *   zf ) \
	(H)+=ds;
	float const tp = two_pi();
	for (int iter = 0; iter < (int)num_iters; ++iter) {
		if (iter > 0 && (unsigned)iter % sweep == 0) { /* end of a sweep: map += deltas */
			for (size_t i = 0; i < (size_t)NX*NY; ++i) {S.mh[i] = S.mh[i] + (float)((double)S.dfix[i]*(1.0/1099511627776.0)); S.dfix[i] = 0;}
		}
		tw_rng rgen; to_rng_set(&rgen, iter+11, 79*iter+121);
		int xi = PAD + (to_rng_rand(&rgen)%xsize);
		int zi = PAD + (to_rng_rand(&rgen)%ysize);
		int const zstart = zi;
		S.have = 0; /* a new droplet knows nothing of the previous one's writes */
		float xp=xi, zp=zi, xf=0, zf=0, s=0, v=0, w=1, dx=0, dz=0;
		float h=sw_read(&S, xi, zi), h00=h, h10=sw_read(&S, xi+1, zi), h01=sw_read(&S, xi, zi+1), h11=sw_read(&S, xi+1, zi+1);
		unsigned numMoves = 0;
		for (; numMoves < MAX_PATH_LEN; ++numMoves) {
			if ((unsigned)(zi - zstart + halo_rule) > 2u*(unsigned)halo_rule) break; /* too far from the start row: the droplet ends */
			++steps;
			{ /* keep the cells this move can touch inside the private view (same policy as droplet_kernel<M_WINDOW>) */
				int const cx = CLAMPI(xi, NX-1), cz = CLAMPI(zi, NY-1);
				int const lx = (cx-2 > 0) ? cx-2 : 0, hx = (cx+3 < NX-1) ? cx+3 : NX-1, lz = (cz-2 > 0) ? cz-2 : 0, hz = (cz+3 < NY-1) ? cz+3 : NY-1;
				int const covered = S.have && lx >= S.wx0 && hx < S.wx0 + S.WX && lz >= S.wz0 && hz < S.wz0 + S.WY;
				if (!covered) {
					float const tx = dx*(float)(S.WX/2 - 5), tz = dz*(float)(S.WY/2 - 5);
					int const bx = (tx != tx) ? 0 : (int)lrintf(tx), bz = (tz != tz) ? 0 : (int)lrintf(tz); /* round to nearest even; NaN -> 0 */
					int ox = cx - S.WX/2 + bx, oz = cz - S.WY/2 + bz;
					if (ox > NX - S.WX) ox = NX - S.WX; if (ox < 0) ox = 0;
					if (oz > NY - S.WY) oz = NY - S.WY; if (oz < 0) oz = 0;
					S.wx0 = ox; S.wz0 = oz;
					for (int r = 0; r < S.WY; ++r) {for (int c = 0; c < S.WX; ++c) {S.win[r*S.WX + c] = S.mh[(size_t)NX*(oz + r) + ox + c];}}
					S.have = 1;
				}
			}
			float gx=h00+h01-h10-h11, gz=h00+h10-h01-h11;
			dx=(dx-gx)*Ki+gx;
			dz=(dz-gz)*Ki+gz;
			float dl=sqrtf(dx*dx+dz*dz);
			if (dl<=FLT_EPSILON) {float a=to_rng_rand_float(&rgen)*tp; dx=cosf(a); dz=sinf(a);}
			else {dx/=dl; dz/=dl;}
			float nxp=xp+dx, nzp=zp+dz;
			if (!(fabsf(nxp) < 2147483648.0f && fabsf(nzp) < 2147483648.0f)) break; /* not a finite in-range position (a NaN in the droplet's own view): the droplet ends; rule of the batched algorithm */
			int nxi=(int)floorf(nxp), nzi=(int)floorf(nzp);
			float nxf=nxp-nxi, nzf=nzp-nzi;
			float nh00=sw_read(&S, nxi, nzi), nh10=sw_read(&S, nxi+1, nzi), nh01=sw_read(&S, nxi, nzi+1), nh11=sw_read(&S, nxi+1, nzi+1);
			float nh=(nh00*(1-nxf)+nh10*nxf)*(1-nzf)+(nh01*(1-nxf)+nh11*nxf)*nzf;
			if (std_max(std_max(nh00, nh10), std_max(nh01, nh11)) < ep->water_plane_z - ep->half_dxy) break;
			int const outside = (xi < 0 || zi < 0 || xi >= NX || zi >= NY);
			if (nh>=h || outside) {
				float ds=(nh-h)+0.001f;
				if (ds>=s || outside) {ds=s; SW_DEPOSIT(h) s=0; break;}
				SW_DEPOSIT(h)
				s-=ds;
				v=0;
			}
			float dh=h-nh;
			float q=std_max(dh, minSlope)*v*w*Kq;
			float ds=s-q;
			if (ds>=0) {ds*=Kd; SW_DEPOSIT(dh) s-=ds;}
			else {
				ds*=-Kr;
				ds=std_min(ds, dh*0.99f);
				{float const relh = ep->relh_adj_tex + (nh - ep->zmin)/(ep->zmax - ep->zmin); ds = (float)(ds*((relh > ep->clip_hd1) ? 0.5 : 2.0));}
				for (int z=zi-1; z<=zi+2; ++z) {
					float zo=z-zp, zo2=zo*zo;
					for (int x=xi-1; x<=xi+2; ++x) {
						float xo=x-xp;
						float wgt=1-(xo*xo+zo2)*0.25f;
						if (wgt<=0) continue;
						wgt*=0.1591549430918953f;
						float const delta=ds*erode_amount*wgt;
						sw_add(&S, CLAMPI(x, NX-1), CLAMPI(z, NY-1), -delta);
					}
				}
				dh-=ds;
				s+=ds;
			}
			v=sqrtf(v*v+Kg*dh);
			w*=1-Kw;
			xp=nxp; zp=nzp; xi=nxi; zi=nzi; xf=nxf; zf=nzf;
			h=nh; h00=nh00; h10=nh10; h01=nh01; h11=nh11;
		}
	}
	for (size_t i = 0; i < (size_t)NX*NY; ++i) {S.mh[i] = S.mh[i] + (float)((double)S.dfix[i]*(1.0/1099511627776.0));}
	for (int y = 0; y < ysize; ++y) {
		for (int x = 0; x < xsize; ++x) {heightmap[(size_t)y*xsize + x] = std_max(min_zval, S.mh[(size_t)(y+PAD)*NX + x+PAD]);}
	}
	free(S.mh); free(S.dfix);
	return steps;
#undef SW_DEPOSIT_AT
#undef SW_DEPOSIT
}
#undef CLAMPI
#undef HMAP_INDEX


/* ------------------------------------------------------------------ per-tile normals and AO (ref: src/tiled_mesh.cpp:586-662,865-880; src/tiled_mesh.h:281-284)
 * tiled_mesh.cpp cannot be linked into oracle/_ref (it pulls in the renderer), so these two loops are restatements; the normal arithmetic
 * is pinned against the reference's own vector3d::get_norm() through oracle/_ref (ref_tile_normals), the AO loop is integer logic on pinned heights. */
void to_tile_normals(const float *zvals_all, unsigned ntiles, unsigned zvsize, float dx_val, float dy_val, unsigned char *rgba, float *min_normal_z) {
	unsigned const stride = zvsize - 1;
	float const dxdy = dx_val*dy_val; /* ref: src/matrix_ops.cpp:80 */
	for (unsigned t = 0; t < ntiles; ++t) {
		const float *zvals = zvals_all + (size_t)t*zvsize*zvsize;
		unsigned char *out = rgba + (size_t)t*stride*stride*4;
		float mnz = 1.0f;
		for (unsigned y = 0; y < stride; ++y) {
			for (unsigned x = 0; x < stride; ++x) {
				unsigned const ix = y*stride + x, ix2 = y*zvsize + x;
				float n[3] = {dy_val*(zvals[ix2] - zvals[ix2 + 1]), dx_val*(zvals[ix2] - zvals[ix2 + zvsize]), dxdy}; /* get_norm_not_normalized */
				float const vmag = sqrtf(n[0]*n[0] + n[1]*n[1] + n[2]*n[2]);
				if (!(vmag < 1.0E-12f)) {n[0] /= vmag; n[1] /= vmag; n[2] /= vmag;} /* pointT::get_norm, src/3DWorld.h:297-300 */
				mnz = std_min(mnz, n[2]);
				for (int i = 0; i < 3; ++i) {out[4*ix + i] = (unsigned char)(127.0*(n[i] + 1.0));}
				out[4*ix + 3] = 0;
			}
		}
		if (min_normal_z) {min_normal_z[t] = mnz;}
	}
}
/* czv_all: ntiles context grids of (stride + 72)^2 heights at origin (x1 - 36, y1 - 36). use_ao_zvals: the GPU-gen-mode flow, where create_zvals
   kept the un-eroded context in ao_zvals and calc_mesh_ao_lighting swaps it in WITHOUT substituting zvals inside the tile (ref: :479-487,604) */
void to_tile_ao(const float *zvals_all, const float *czv_all, unsigned ntiles, unsigned zvsize, float half_dxy, int use_ao_zvals, unsigned char *ao) {
	enum {NUM_AO_DIRS = 8, NUM_AO_STEPS = 8, AO_RAY_LEN = 36};
	unsigned const stride = zvsize - 1, context_sz = stride + 2*AO_RAY_LEN;
	int dirs[NUM_AO_DIRS][2], ix = 0;
	for (int y = -1; y <= 1; ++y) {for (int x = -1; x <= 1; ++x) {if (x != 0 || y != 0) {dirs[ix][0] = x; dirs[ix][1] = y; ++ix;}}}
	float const dz = (float)(0.5*half_dxy);
	for (unsigned t = 0; t < ntiles; ++t) {
		const float *zvals = zvals_all + (size_t)t*zvsize*zvsize, *gen = czv_all + (size_t)t*context_sz*context_sz;
		float *czv = (float *)malloc((size_t)context_sz*context_sz*sizeof(float));
		for (int y = 0; y < (int)context_sz; ++y) { /* ref: :624-637 */
			for (int x = 0; x < (int)context_sz; ++x) {
				int const xv = x - AO_RAY_LEN, yv = y - AO_RAY_LEN;
				czv[y*context_sz + x] = (!use_ao_zvals && xv >= 0 && yv >= 0 && xv < (int)zvsize && yv < (int)zvsize) ? zvals[yv*zvsize + xv] : gen[y*context_sz + x];
			}
		}
		for (int y = 0; y < (int)stride; ++y) { /* ref: :640-659 */
			for (int x = 0; x < (int)stride; ++x) {
				unsigned atten = 0;
				for (unsigned d = 0; d < NUM_AO_DIRS; ++d) {
					float z0 = zvals[y*zvsize + x];
					int sx = dirs[d][0], sy = dirs[d][1], vx = x, vy = y;
					for (unsigned s = 0; s < NUM_AO_STEPS; ++s) {
						vx += sx; vy += sy;
						z0 += dz;
						sx += dirs[d][0]; sy += dirs[d][1];
						if (czv[(vy + AO_RAY_LEN)*context_sz + (vx + AO_RAY_LEN)] > z0) {atten += (NUM_AO_STEPS - s); break;}
					}
				}
				float const ao_scale = 1.0 - (float)atten/(float)(NUM_AO_DIRS*NUM_AO_STEPS);
				ao[(size_t)t*stride*stride + y*stride + x] = (unsigned char)(255.0*ao_scale);
			}
		}
		free(czv);
	}
}

/* ------------------------------------------------------------------ heightmap-texture tiles (ref: src/heightmap.cpp:74-77,309-341,385-402; src/mesh_gen.cpp:120) */
static int round_fp_f(float v) {return (v > 0.0f) ? (int)(v + 0.5f) : (int)(v - 0.5f);} /* ref: src/inlines.h:63 */
static int hmap_clamp_no_scale(int *x, int *y, const tw_hmap_sampler *H) { /* ref: src/heightmap.cpp:315-341 */
	*x += H->width/2; *y += H->height/2;
	if (*x >= 0 && *y >= 0 && *x < H->width && *y < H->height) return 1;
	switch (H->edge_mode) {
	case 0:
		*x = (0 < ((H->width -1 < *x) ? H->width -1 : *x)) ? ((H->width -1 < *x) ? H->width -1 : *x) : 0;
		*y = (0 < ((H->height-1 < *y) ? H->height-1 : *y)) ? ((H->height-1 < *y) ? H->height-1 : *y) : 0;
		break;
	case 1: return 0;
	default: {
		int const xmod = abs(*x)%H->width, ymod = abs(*y)%H->height, xdiv = *x/H->width, ydiv = *y/H->height;
		*x = (xdiv & 1) ? (H->width  - xmod - 1) : xmod;
		*y = (ydiv & 1) ? (H->height - ymod - 1) : ymod;
		}
	}
	return 1;
}
static float hmap_scale_val(float val, const tw_hmap_sampler *H) {return (H->h_scale*H->mesh_file_scale*val + H->mesh_file_tz)*H->mesh_scale_z_inv;}
static float hmap_raw_height(const unsigned char *data, int x, int y, const tw_hmap_sampler *H) {
	size_t const ix = (size_t)H->width*y + x;
	float const v = (float)(data[ix<<1]/256.0 + data[(ix<<1)+1]); /* get_heightmap_value, ncolors == 2 */
	return hmap_scale_val(v, H);
}
void to_hmap_sample_tiles(const unsigned char *data16, const tw_hmap_sampler *H, const int *origins_xy, unsigned ntiles, unsigned zvsize, float *out) {
	for (unsigned t = 0; t < ntiles; ++t) {
		for (unsigned yy = 0; yy < zvsize; ++yy) {
			for (unsigned xx = 0; xx < zvsize; ++xx) {
				int x = origins_xy[2*t] + (int)xx, y = origins_xy[2*t+1] + (int)yy;
				float z;
				if (H->mesh_scale < 1.0) { /* interpolate_height */
					float const sx = H->mesh_scale*(float)x, sy = H->mesh_scale*(float)y;
					int xlo = (int)floorf(sx), ylo = (int)floorf(sy), xhi = (int)ceilf(sx), yhi = (int)ceilf(sy);
					float const xv = sx - xlo, yv = sy - ylo;
					if (!hmap_clamp_no_scale(&xlo, &ylo, H) || !hmap_clamp_no_scale(&xhi, &yhi, H)) {z = hmap_scale_val(0.0f, H);}
					else {
						z = yv*(xv*hmap_raw_height(data16, xhi, yhi, H) + (1.0f-xv)*hmap_raw_height(data16, xlo, yhi, H)) +
						    (1.0f-yv)*(xv*hmap_raw_height(data16, xhi, ylo, H) + (1.0f-xv)*hmap_raw_height(data16, xlo, ylo, H));
					}
				}
				else {
					x = round_fp_f(H->mesh_scale*(x + 0.0f)); y = round_fp_f(H->mesh_scale*(y + 0.0f));
					z = hmap_clamp_no_scale(&x, &y, H) ? hmap_raw_height(data16, x, y, H) : hmap_scale_val(0.0f, H);
				}
				out[((size_t)t*zvsize + yy)*zvsize + xx] = z;
			}
		}
	}
}

/* ------------------------------------------------------------------ point queries (ref: src/mesh_gen.cpp:797-847) */
static float eval_mesh_sin_terms_scaled(float xval, float yval, float xy_scale, int MX, int MY, const tw_height_params *p, const float *tab, const float *T) { /* ref: :807-813 */
	float const xv = xy_scale*(xval - (float)(MX >> 1)), yv = xy_scale*(yval - (float)(MY >> 1));
	if (p->gen_mode != TW_MGEN_SINE) {return to_get_noise_zval(xv, yv, p);}
	float val = to_eval_mesh_sin_terms(p->mesh_scale*xv, p->mesh_scale*yv, tab, T, p->start_eval_sin)*p->mesh_scale_z_inv;
	apply_noise_shape_final(&val, p->gen_shape, &p->hmap);
	return val;
}
void to_eval_points(const float *xy, size_t n, const tw_height_params *p, const tw_point_query *q, const float *tab, const float *T, float *out) {
	for (size_t i = 0; i < n; ++i) {
		float const xin = xy[2*i], yin = xy[2*i + 1];
		if (q->kind == TW_PQ_SIN_TERMS) {out[i] = to_eval_mesh_sin_terms(xin, yin, tab, T, p->start_eval_sin); continue;}
		if (q->kind == TW_PQ_SIN_TERMS_SCALED) {out[i] = eval_mesh_sin_terms_scaled(xin, yin, q->xy_scale, q->mesh_x_size, q->mesh_y_size, p, tab, T); continue;}
		/* get_exact_zval, procedural branch (ref: :816-847) */
		float xval = (float)((xin + q->x_scene_size)*p->dx_val_inv + 0.5);
		float yval = (float)((yin + q->y_scene_size)*p->dy_val_inv + 0.5);
		if (!q->no_xyoff) {xval += q->xoff2; yval += q->yoff2;}
		float zval = eval_mesh_sin_terms_scaled(xval, yval, 1.0f, q->mesh_x_size, q->mesh_y_size, p, tab, T);
		if (p->glaciate) { /* apply_glaciate, ref: :380-385 */
			float const zmax_est = p->zmax_est, zmax_est2 = (float)(2.0*zmax_est), zmax_est2_inv = (float)(1.0/zmax_est2);
			float const relh = (zval + zmax_est)*zmax_est2_inv;
			zval = do_glaciate_exp(relh, p->custom_glaciate_exp)*zmax_est2 - zmax_est;
		}
		apply_mesh_sine(&zval, (xval - (float)(q->mesh_x_size >> 1)), (yval - (float)(q->mesh_y_size >> 1)), p, tab);
		out[i] = zval;
	}
}

/* ------------------------------------------------------------------ 3-D noise / voxels */
void to_noise3d_gen_sines(int rs1, int rs2, float mag, float freq, float *rdata) { /* ref: src/upsurface.cpp:16-38 */
	tw_rng r; to_rng_set(&r, rs1, rs2);
	float const tp = two_pi();
	for (unsigned i = 0; i < 5; ++i) {
		for (unsigned j = 0; j < 12; ++j) {
			float *e = rdata + 7*(12*i + j);
			e[0] = to_rng_rand_uniform(&r, 0.2f, 1.0f)*mag;
			e[1] = to_rng_rand_uniform(&r, 0.1f, 1.0f)*freq;
			e[2] = (float)(to_rng_randd(&r)*tp);
			e[3] = to_rng_rand_uniform(&r, 0.1f, 1.0f)*freq;
			e[4] = (float)(to_rng_randd(&r)*tp);
			e[5] = to_rng_rand_uniform(&r, 0.1f, 1.0f)*freq;
			e[6] = (float)(to_rng_randd(&r)*tp);
		}
		mag  *= 0.5f;
		freq /= 0.4f;
	}
}
float to_noise3d_get_val_pt(const float *rdata, const float *tab, float px, float py, float pz) { /* ref: src/upsurface.cpp:73-85 */
	float val = 0.0f;
	for (unsigned k = 0; k < 60; ++k) {
		float const *e = rdata + 7*k;
		float const x = to_sinf_lut(tab, e[1]*px + e[2]);
		float const y = to_sinf_lut(tab, e[3]*py + e[4]);
		float const z = to_sinf_lut(tab, e[5]*pz + e[6]);
		val += e[0]*x*y*z;
	}
	return val;
}

static inline float clip_pm1(float x) {return std_max(-1.0f, std_min(1.0f, x));} /* CLIP_TO_pm1, ref: src/3DWorld.h:149 */

void to_voxel_fill(const tw_voxel_params *vp, const float *rdata_in, const float *tab, float *out, int nthreads)
{
	unsigned const nx = vp->nx, ny = vp->ny, nz = vp->nz, num[3] = {nx, ny, nz};
	float rdata[420];
	float *xyz[3] = {NULL, NULL, NULL};
	if (vp->gen_mode == TW_MGEN_SINE) {
		if (rdata_in) {memcpy(rdata, rdata_in, sizeof(rdata));} else {to_noise3d_gen_sines(vp->rseed1, vp->rseed2, vp->mag, vp->freq, rdata);}
		for (unsigned d = 0; d < 3; ++d) { /* gen_xyz_vals, ref: src/upsurface.cpp:41-57 */
			xyz[d] = (float *)malloc((size_t)60*num[d]*sizeof(float));
			float val = vp->lo_pos[d] + vp->offset[d];
			for (unsigned i = 0; i < num[d]; ++i) {
				for (unsigned k = 0; k < 60; ++k) {
					unsigned const index2 = 7*k + 2*d;
					float v = to_sinf_lut(tab, rdata[index2+1]*val + rdata[index2+2]);
					if (d == 0) {v *= rdata[index2];}
					xyz[d][(size_t)i*60 + k] = v;
				}
				val += vp->vsz[d];
			}
		}
	}
#ifdef _OPENMP
	if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
	(void)nthreads;
	float const rx = vp->rx, ry = vp->ry;
#pragma omp parallel for schedule(static,1) num_threads(nthreads)
	for (int y = 0; y < (int)ny; ++y) { /* ref: src/voxels.cpp:312-345 */
		for (unsigned x = 0; x < nx; ++x) {
			for (unsigned z = 0; z < nz; ++z) {
				float val = 0.0f;
				if (vp->gen_mode == TW_MGEN_SINE) { /* get_val(x,y,z,tables), ref: src/upsurface.cpp:60-70 */
					float const *xv = xyz[0] + (size_t)x*60, *yv = xyz[1] + (size_t)y*60, *zv = xyz[2] + (size_t)z*60;
					for (unsigned k = 0; k < 60; ++k) {val += xv[k]*yv[k]*zv[k];}
				}
				else {
					float const px = ((float)x*vp->vsz[0] + vp->lo_pos[0]) + vp->offset[0];
					float const py = ((float)y*vp->vsz[1] + vp->lo_pos[1]) + vp->offset[1];
					float const pz = ((float)z*vp->vsz[2] + vp->lo_pos[2]) + vp->offset[2];
					float nmag = vp->mag, nfreq = (float)(0.25*vp->freq);
					float const lacunarity = 1.92f, gain = 0.5f;
					for (int n = 0; n < vp->octaves; ++n) {
						float const nvx = nfreq*px + rx, nvy = nfreq*py + ry, nvz = nfreq*pz + (rx - ry);
						val   += nmag*((vp->gen_mode == TW_MGEN_PERLIN) ? to_perlin3(nvx, nvy, nvz) : to_simplex3(nvx, nvy, nvz));
						nmag  *= gain;
						nfreq *= lacunarity;
					}
				}
				val += (float)z*vp->zscale;
				if (vp->normalize_to_1) {val = clip_pm1(val);}
				out[z + ((size_t)x + (size_t)y*nx)*nz] = val;
			}
		}
	}
	for (int d = 0; d < 3; ++d) {free(xyz[d]);}
	/* optional attenuation passes, ref: src/voxels.cpp:403-482 */
	float const aval = vp->atten_val;
	if (vp->atten_mode == 2) { /* atten_at_edges */
#pragma omp parallel for schedule(static) num_threads(nthreads)
		for (int y = 0; y < (int)ny; ++y) {
			float const vy = (float)(1.0 - 2.0*fabs(y - 0.5*ny)/(float)ny);
			for (unsigned x = 0; x < nx; ++x) {
				float const vx = (float)(1.0 - 2.0*fabs(x - 0.5*nx)/(float)nx);
				for (unsigned z = 0; z < nz; ++z) {
					float const vz = (float)(1.0 - 2.0*fabs(z - 0.5*nz)/(float)nz), v = 0.25f - vx*vy*vz;
					if (v > 0.0) {float *o = out + z + ((size_t)x + (size_t)y*nx)*nz; *o = (float)(*o + 8.0*aval*v);}
				}
			}
		}
	}
	else if (vp->atten_mode == 1) { /* atten_at_top_only, atten_top_mode 0 */
#pragma omp parallel for schedule(static) num_threads(nthreads)
		for (int y = 0; y < (int)ny; ++y) {
			for (unsigned x = 0; x < nx; ++x) {
				for (unsigned z = 0; z < nz; ++z) {
					float const z_atten = (float)(z/(float)nz - 0.75);
					if (z_atten > 0.0) {float *o = out + z + ((size_t)x + (size_t)y*nx)*nz; *o += aval*z_atten;}
				}
			}
		}
	}
	else if (vp->atten_mode >= 3 && vp->atten_mode <= 5) { /* atten_to_sphere(val, inner_radius, atten_inner=(mode>=4), no_atten_zbot=(mode==5)) */
		float const two_nz_inv = (float)(2.0/(float)nz), inner_radius = vp->atten_inner_radius;
		int const atten_inner = (vp->atten_mode >= 4), no_atten_zbot = (vp->atten_mode == 5);
#pragma omp parallel for schedule(static) num_threads(nthreads)
		for (int y = 0; y < (int)ny; ++y) {
			float const vy = (float)(2.0*fabs(y - 0.5*ny)/(float)ny);
			for (unsigned x = 0; x < nx; ++x) {
				float const vx = (float)(2.0*fabs(x - 0.5*nx)/(float)nx);
				for (unsigned z = 0; z < nz; ++z) {
					float const deltaz = (float)(z - 0.5*nz), zval = no_atten_zbot ? std_max(0.0f, deltaz) : fabsf(deltaz);
					float const vz = zval*two_nz_inv, radius = sqrtf(vx*vx + vy*vy + vz*vz);
					float adj = 0.0f;
					if (radius > inner_radius) {adj = (radius - inner_radius)/(1.0f - inner_radius);}
					else if (atten_inner) {adj = (radius - inner_radius)/inner_radius;}
					out[z + ((size_t)x + (size_t)y*nx)*nz] += aval*adj;
				}
			}
		}
	}
}

/* ------------------------------------------------------------------ heightmap 16-bit pack (ref: src/heightmap.cpp:191-215, src/Textures.cpp:1889-1893) */
size_t to_from_floats_u16(const float *vals, size_t n, float val_mult, float val_add, unsigned char *out) {
	float const val_div = (float)(1.0/val_mult);
	size_t bad = 0;
	for (size_t i = 0; i < n; ++i) {
		float const v = (vals[i] - val_add)*val_div;
		if (!(v >= 0.0 && v < 256.0)) {++bad; out[2*i] = out[2*i+1] = 0; continue;}
		unsigned char const high_bits = (unsigned char)v;
		out[2*i+1] = high_bits;
		out[2*i]   = (unsigned char)(256.0f*(v - (float)high_bits));
	}
	return bad;
}
void to_to_floats_u16(const unsigned char *data, size_t n, float val_mult, float val_add, float *vals) {
	for (size_t i = 0; i < n; ++i) {
		float v = (float)(data[2*i]/256.0 + data[2*i+1]);
		vals[i] = val_mult*v + val_add;
	}
}


/* ------------------------------------------------------------------ voxel post-processing
 * Pinned against the reference's own voxel_manager member functions, cut out of src/voxels.cpp at build time (oracle/refbuild/build_ref.sh,
 * tests/test_oracle_vs_reference.py). */
static int val_is_outside(float val, float isolevel, int invert) { /* ref: src/voxels.cpp:571-574 */
	return (val == isolevel) ? 1 : ((((val < isolevel) ? 1 : 0) ^ (invert ? 1 : 0)) ? 1 : 0);
}
void to_voxel_outside(const float *vals, const tw_voxel_post_params *vp, const unsigned *zix_xy, unsigned char *outside) { /* ref: :577-604 */
	unsigned const nx = vp->nx, ny = vp->ny, nz = vp->nz;
	for (unsigned y = 0; y < ny; ++y) {
		for (unsigned x = 0; x < nx; ++x) {
			unsigned const zix = zix_xy ? zix_xy[y*nx + x] : 0;
			for (unsigned z = 0; z < nz; ++z) {
				size_t const i = z + ((size_t)x + (size_t)y*nx)*nz;
				int const on_edge = (vp->make_closed_surface && ((x == 0 || x == nx-1) || (y == 0 || y == ny-1) || (z == 0 || z == nz-1)));
				unsigned char ival = on_edge ? TW_VOX_ON_EDGE : (unsigned char)val_is_outside(vals[i], vp->isolevel, vp->invert);
				if (z < zix) {ival |= TW_VOX_UNDER_MESH;}
				outside[i] = ival;
			}
		}
	}
}
/* flood_fill_range (ref: :729-757) over the whole grid with an explicit stack */
static void flood_fill(unsigned char *outside, unsigned nx, unsigned ny, unsigned nz, unsigned *work, size_t nwork, unsigned char fill_val, unsigned char bit_mask) {
	unsigned const nxnz = nx*nz;
	while (nwork) {
		unsigned const cur = work[--nwork];
		unsigned const y = cur/nxnz, cur_xz = cur - y*nxnz, x = cur_xz/nz, z = cur_xz - x*nz;
#define FF_INNER(pos, max_range, step) \
		if (pos >= 1)            {unsigned const ix = cur - step; if (outside[ix] == fill_val) {work[nwork++] = ix; outside[ix] |= bit_mask;}} \
		if (pos + 1 < max_range) {unsigned const ix = cur + step; if (outside[ix] == fill_val) {work[nwork++] = ix; outside[ix] |= bit_mask;}}
		FF_INNER(x, nx, nz)
		FF_INNER(y, ny, nxnz)
		FF_INNER(z, nz, 1)
#undef FF_INNER
	}
}
unsigned long long to_voxel_remove_unconnected(float *vals, unsigned char *outside, const tw_voxel_post_params *vp) {
	unsigned const nx = vp->nx, ny = vp->ny, nz = vp->nz;
	size_t const n = (size_t)nx*ny*nz;
	unsigned long long changed = 0;
	float const TOLERANCE = 1.0E-12f;
	if (vp->remove_unconnected <= 0) return 0;
	unsigned *work = (unsigned *)malloc((n + 1)*sizeof(unsigned)*2); /* a voxel is pushed at most once by the fill, seeds may be pushed twice */
	size_t nwork = 0;
	/* remove_unconnected_outside_range(keep_at_edge, 0, 0, nx, ny), ref: :759-827 */
	if (vp->centre_seed) {
		size_t const ix = (nz/2) + ((size_t)(nx/2) + (size_t)(ny/2)*nx)*nz;
		work[nwork++] = (unsigned)ix; outside[ix] |= TW_VOX_ANCHORED;
	}
	else {
		for (size_t ix = 0; ix < n; ++ix) {if (outside[ix] == TW_VOX_UNDER_MESH) {work[nwork++] = (unsigned)ix; outside[ix] |= TW_VOX_ANCHORED;}}
	}
	if (vp->keep_at_edge) {
		for (unsigned y = 0; y < ny; ++y) {
			for (unsigned x = 0; x < nx; ++x) {
				if (x != 0 && x+1 != nx && y != 0 && y+1 != ny) continue;
				for (unsigned z = 0; z < nz; ++z) {
					size_t const ix = z + ((size_t)x + (size_t)y*nx)*nz;
					if (outside[ix] == 1) continue;
					if (!(outside[ix] & TW_VOX_ANCHORED)) {work[nwork++] = (unsigned)ix;} /* the reference pushes duplicates too; harmless either way */
					outside[ix] |= TW_VOX_ANCHORED;
				}
			}
		}
	}
	flood_fill(outside, nx, ny, nz, work, nwork, 0, TW_VOX_ANCHORED);
	for (size_t ix = 0; ix < n; ++ix) {
		if (outside[ix] > 1) {outside[ix] &= (unsigned char)~TW_VOX_ANCHORED;}
		else if (outside[ix] != 1) { /* inside and not anchored: make_voxel_outside, ref: :861-864 */
			outside[ix] = 1;
			vals[ix] = vp->isolevel - (vp->invert ? -TOLERANCE : TOLERANCE);
			++changed;
		}
	}
	if (vp->remove_unconnected > 2) { /* remove_interior_holes, ref: :831-858 */
		nwork = 0;
		for (unsigned y = 0; y < ny; ++y) {
			for (unsigned x = 0; x < nx; ++x) {
				size_t const ix = (nz-1) + ((size_t)x + (size_t)y*nx)*nz;
				if (outside[ix]) {work[nwork++] = (unsigned)ix; outside[ix] |= TW_VOX_ANCHORED;}
			}
		}
		if (nwork) {
			flood_fill(outside, nx, ny, nz, work, nwork, 1, TW_VOX_ANCHORED);
			for (size_t ix = 0; ix < n; ++ix) {
				if (outside[ix] & TW_VOX_ANCHORED) {outside[ix] &= (unsigned char)~TW_VOX_ANCHORED;}
				else if (outside[ix] == 1) { /* make_voxel_inside, ref: :865-868 */
					outside[ix] = 0;
					vals[ix] = vp->isolevel + (vp->invert ? -TOLERANCE : TOLERANCE);
					++changed;
				}
			}
		}
	}
	free(work);
	return changed;
}
/* interpolate_pt, ref: :485-493 */
static void interpolate_pt(float isolevel, const float *pt1, const float *pt2, float val1, float val2, float *pt) {
	float const TOLERANCE = 1.0E-12f;
	if (fabsf(isolevel - val1) < TOLERANCE) {pt[0] = pt1[0]; pt[1] = pt1[1]; pt[2] = pt1[2]; return;}
	if (fabsf(isolevel - val2) < TOLERANCE) {pt[0] = pt2[0]; pt[1] = pt2[1]; pt[2] = pt2[2]; return;}
	if (fabsf(val1     - val2) < TOLERANCE) {pt[0] = pt1[0]; pt[1] = pt1[1]; pt[2] = pt1[2]; return;}
	float const mu = std_max(0.0f, std_min(1.0f, (isolevel - val1)/(val2 - val1))); /* CLIP_TO_01 */
	for (int i = 0; i < 3; ++i) {pt[i] = pt1[i] + mu*(pt2[i] - pt1[i]);}
}
/* add_triangles_for_voxel at lod 0 for every cube in (y, x, z) order, unwelded (ref: :495-566); returns the number of triangles produced */
unsigned long long to_voxel_triangles(const float *vals, const unsigned char *outside, const tw_voxel_post_params *vp, const unsigned *edge_table, const int *tri_table,
                                      const unsigned *edge_to_vals, float *tris, unsigned long long capacity)
{
	unsigned const nx = vp->nx, ny = vp->ny, nz = vp->nz;
	unsigned long long n = 0;
	for (unsigned y = 0; y < ny; ++y) {
		for (unsigned x = 0; x < nx; ++x) {
			for (unsigned z = 0; z < nz; ++z) {
				unsigned const x2 = (x+1 < nx-1) ? x+1 : nx-1, y2 = (y+1 < ny-1) ? y+1 : ny-1, z2 = (z+1 < nz-1) ? z+1 : nz-1;
				unsigned const xv[2] = {x, x2}, yv[2] = {y, y2}, zv[2] = {z, z2};
				if (x2 <= x || y2 <= y || z2 <= z) continue;
				unsigned cix = 0;
				int all_under_mesh = (vp->skip_under_mesh != 0);
				for (unsigned yhi = 0; yhi < 2; ++yhi) {
					for (unsigned xhi = 0; xhi < 2; ++xhi) {
						size_t const ix = z + ((size_t)xv[xhi] + (size_t)yv[yhi]*nx)*nz;
						if (all_under_mesh) {all_under_mesh = ((outside[ix] & TW_VOX_UNDER_MESH) != 0);}
						for (unsigned zhi = 0; zhi < 2; ++zhi) {if (outside[ix + zv[zhi]-z] & 7) {cix |= 1u << ((xhi^yhi) + 2*yhi + 4*zhi);}}
					}
				}
				if (all_under_mesh) continue;
				unsigned const edge_val = edge_table[cix];
				if (edge_val == 0) continue;
				const int *t = tri_table + 16*cix;
				float const cube[3][2] = {{x*vp->vsz[0] + vp->lo_pos[0], x2*vp->vsz[0] + vp->lo_pos[0]}, {y*vp->vsz[1] + vp->lo_pos[1], y2*vp->vsz[1] + vp->lo_pos[1]},
				                          {z*vp->vsz[2] + vp->lo_pos[2], z2*vp->vsz[2] + vp->lo_pos[2]}};
				float vlist[12][3];
				for (unsigned i = 0; i < 12; ++i) {
					if (!(edge_val & (1u << i))) continue;
					float v2[2], pts[2][3];
					for (unsigned d = 0; d < 2; ++d) {
						unsigned const e = edge_to_vals[2*i + d], yhi = (e & 2) >> 1, xhi = yhi ^ (e & 1), zhi = e >> 2;
						size_t const ix = zv[zhi] + ((size_t)xv[xhi] + (size_t)yv[yhi]*nx)*nz;
						v2[d] = ((outside[ix] & 7) == TW_VOX_ON_EDGE) ? vp->isolevel : vals[ix];
						pts[d][0] = cube[0][xhi]; pts[d][1] = cube[1][yhi]; pts[d][2] = cube[2][zhi];
					}
					interpolate_pt(vp->isolevel, pts[0], pts[1], v2[0], v2[1], vlist[i]);
				}
				for (unsigned i = 0; t[i] >= 0; i += 3) {
					const float *p0 = vlist[t[i]], *p1 = vlist[t[i+1]], *p2 = vlist[t[i+2]];
					float const a[3] = {p1[0]-p0[0], p1[1]-p0[1], p1[2]-p0[2]}, b[3] = {p2[0]-p1[0], p2[1]-p1[1], p2[2]-p1[2]}; /* get_normal: cross(v2 - v1, v3 - v2) */
					float const cx = a[1]*b[2] - a[2]*b[1], cy = a[2]*b[0] - a[0]*b[2], cz = a[0]*b[1] - a[1]*b[0];
					if (cx == 0.0f && cy == 0.0f && cz == 0.0f) continue; /* normalize() leaves a zero vector zero: "invalid triangle", ref: :550 */
					if (n < capacity) {float *o = tris + 9*n; for (int k = 0; k < 3; ++k) {o[k] = p0[k]; o[3+k] = p1[k]; o[6+k] = p2[k];}}
					++n;
				}
			}
		}
	}
	return n;
}


/* ------------------------------------------------------------------ mesh shadows
 * Pinned against the reference's own mesh_shadow_gen / calc_mesh_shadows / do_line_clip, cut out of src/visibility.cpp and src/Math3d.cpp at build time. */
typedef struct {float x, y, z;} sh_pt;
static int sh_region(sh_pt v, float d[3][2]) { /* get_region, ref: src/inlines.h:522-528 */
	int region = 0;
	if (v.x < d[0][0]) {region |= 0x01;} else if (v.x >= d[0][1]) {region |= 0x02;}
	if (v.y < d[1][0]) {region |= 0x04;} else if (v.y >= d[1][1]) {region |= 0x08;}
	if (v.z < d[2][0]) {region |= 0x10;} else if (v.z >= d[2][1]) {region |= 0x20;}
	return region;
}
static int sh_line_clip(sh_pt *v1, sh_pt *v2, float d[3][2]) { /* do_line_clip, ref: src/Math3d.cpp:1029-1034,1070-1086 */
	float const TOLERANCE = 1.0E-12f;
	int const region1 = sh_region(*v1, d), region2 = sh_region(*v2, d);
	if (region1 & region2) return 0;
	int const region3 = region1 | region2;
	if (region3 == 0) return 1;
	float tmin = 0.0f, tmax = 1.0f;
	sh_pt const dv = {v2->x - v1->x, v2->y - v1->y, v2->z - v1->z};
#define SH_CLIP(reg, va, vb, vd, vc) if (region3 & (reg)) {float const t = ((va) - (vb))/(vd); if ((vc) > 0.0) {if (t > tmin) tmin = t;} else {if (t < tmax) tmax = t;} if (tmin >= tmax) return 0;}
	SH_CLIP(0x01, d[0][0], v1->x, dv.x,  dv.x)
	SH_CLIP(0x02, d[0][1], v1->x, dv.x, -dv.x)
	SH_CLIP(0x04, d[1][0], v1->y, dv.y,  dv.y)
	SH_CLIP(0x08, d[1][1], v1->y, dv.y, -dv.y)
	SH_CLIP(0x10, d[2][0], v1->z, dv.z,  dv.z)
	SH_CLIP(0x20, d[2][1], v1->z, dv.z, -dv.z)
#undef SH_CLIP
	if (tmax > TOLERANCE) {v2->x = v1->x + dv.x*tmax; v2->y = v1->y + dv.y*tmax; v2->z = v1->z + dv.z*tmax;}
	if ((double)tmin < (1.0 - TOLERANCE)) {v1->x += dv.x*tmin; v1->y += dv.y*tmin; v1->z += dv.z*tmin;}
	return 1;
}
typedef struct {const tw_shadow_params *sp; const float *mh, *sh_in_x, *sh_in_y; float *sh_out_x, *sh_out_y; unsigned char *smask; int xsize, ysize; sh_pt dir; float dist;} sh_gen;
static void sh_trace(const sh_gen *G, sh_pt v1) { /* mesh_shadow_gen::trace_shadow_path, ref: src/visibility.cpp:421-477 */
	const tw_shadow_params *sp = G->sp;
	int const xsize = G->xsize, ysize = G->ysize;
	sh_pt v2 = {v1.x + G->dir.x*G->dist, v1.y + G->dir.y*G->dist, v1.z + 0.0f};
	float d[3][2] = {{-sp->x_scene_size, -sp->x_scene_size + sp->dx_val*xsize}, {-sp->y_scene_size, -sp->y_scene_size + sp->dy_val*ysize}, {sp->zmin, sp->zmax}};
	if (!sh_line_clip(&v1, &v2, d)) return;
	int const xa = (int)((v1.x + sp->x_scene_size)*sp->dx_val_inv + 0.5), ya = (int)((v1.y + sp->y_scene_size)*sp->dy_val_inv + 0.5);
	int const xb = (int)((v2.x + sp->x_scene_size)*sp->dx_val_inv + 0.5), yb = (int)((v2.y + sp->y_scene_size)*sp->dy_val_inv + 0.5);
	int const dx = xb - xa, dy = yb - ya;
	int const dim = (fabsf(G->dir.x) < fabsf(G->dir.y));
	double const dir_ratio = G->dir.z/(dim ? G->dir.y : G->dir.x);
	int inited = 0;
	sh_pt cur = {0, 0, 0};
	int x = xa, y = ya, dx1 = 0, dy1 = 0, dx2 = 0, dy2 = 0;
	if (dx < 0) {dx1 = -1; dx2 = -1;} else if (dx > 0) {dx1 = 1; dx2 = 1;}
	if (dy < 0) {dy1 = -1;} else if (dy > 0) {dy1 = 1;}
	int longest = abs(dx), shortest = abs(dy);
	if (longest <= shortest) {
		int const t = longest; longest = shortest; shortest = t;
		if (dy < 0) {dy2 = -1;} else if (dy > 0) {dy2 = 1;}
		dx2 = 0;
	}
	int numerator = longest >> 1;
	for (int i = 0; i <= longest; i++) {
		if (x >= 0 && y >= 0 && x < xsize && y < ysize) {
			sh_pt const pt = {-sp->x_scene_size + sp->dx_val*x, -sp->y_scene_size + sp->dy_val*y, G->mh[y*xsize + x]};
			if (G->sh_in_y != NULL && x == xa && G->sh_in_y[y] > TW_MESH_MIN_Z) {cur.x = pt.x; cur.y = pt.y; cur.z = G->sh_in_y[y]; inited = 1;}
			else if (G->sh_in_x != NULL && y == ya && G->sh_in_x[x] > TW_MESH_MIN_Z) {cur.x = pt.x; cur.y = pt.y; cur.z = G->sh_in_x[x]; inited = 1;}
			float const shadow_z = (float)(((dim ? pt.y : pt.x) - (dim ? cur.y : cur.x))*dir_ratio + cur.z);
			if (inited && shadow_z > pt.z) {
				G->smask[y*xsize + x] |= TW_MESH_SHADOW;
				if (G->sh_out_y != NULL && x == xb) {G->sh_out_y[y] = shadow_z;}
				if (G->sh_out_x != NULL && y == yb) {G->sh_out_x[x] = shadow_z;}
			}
			else {cur = pt;}
			inited = 1;
		}
		numerator += shortest;
		if (numerator >= longest) {numerator -= longest; x += dx1; y += dy1;} else {x += dx2; y += dy2;}
	}
}
void to_calc_mesh_shadows(const tw_shadow_params *sp, const float *mh, unsigned char *smask, int xsize, int ysize, const float *sh_in_x, const float *sh_in_y,
                          float *sh_out_x, float *sh_out_y)
{ /* calc_mesh_shadows + mesh_shadow_gen::run, ref: src/visibility.cpp:478-517 (run_x then run_y: the 1-thread order of the two OpenMP sections) */
	float const TOLERANCE = 1.0E-12f;
	int const all_shadowed = (!sp->no_shadow && sp->lpos[2] < sp->zmin);
	for (int i = 0; i < xsize*ysize; ++i) {smask[i] = all_shadowed ? TW_MESH_SHADOW : 0;}
	if (sp->no_shadow) return;
	if (sp->lpos[0] == 0.0 && sp->lpos[1] == 0.0) return;
	sh_gen G = {sp, mh, sh_in_x, sh_in_y, sh_out_x, sh_out_y, smask, xsize, ysize, {0, 0, 0}, 0.0f};
	float const vmag = sqrtf(sp->lpos[0]*sp->lpos[0] + sp->lpos[1]*sp->lpos[1] + sp->lpos[2]*sp->lpos[2]);
	sh_pt n = {sp->lpos[0], sp->lpos[1], sp->lpos[2]};
	if (!(vmag < TOLERANCE)) {n.x = sp->lpos[0]/vmag; n.y = sp->lpos[1]/vmag; n.z = sp->lpos[2]/vmag;} /* get_norm */
	G.dir.x = -n.x; G.dir.y = -n.y; G.dir.z = -n.z;
	G.dist = (float)(2.0*sp->xy_sum_size/sqrtf(G.dir.x*G.dir.x + G.dir.y*G.dir.y));
	{
		float const xval = -sp->x_scene_size + sp->dx_val*((G.dir.x > 0) ? 0 : xsize);
		for (int y = 0; y < 2*ysize; ++y) {sh_pt const v = {xval, (float)(-sp->y_scene_size + 0.5*sp->dy_val*y), 0.0f}; sh_trace(&G, v);}
	}
	{
		float const yval = -sp->y_scene_size + sp->dy_val*((G.dir.y > 0) ? 0 : ysize);
		for (int x = 0; x < 2*xsize; ++x) {sh_pt const v = {(float)(-sp->x_scene_size + 0.5*sp->dx_val*x), yval, 0.0f}; sh_trace(&G, v);}
	}
}
/* tile_t::calc_shadows_for_light for a batch (ref: src/tiled_mesh.cpp:664-692): a tile takes sh_in from the sh_out of its neighbours toward the light when they are in
 * the batch (d = 0: x neighbour's sh_out[1] -> sh_in_y; d = 1: y neighbour's sh_out[0] -> sh_in_x); tiles are processed so that neighbours come first */
void to_tile_shadows_batch(const float *zvals, const int *tile_xy, unsigned ntiles, unsigned zvsize, const tw_shadow_params *sp, unsigned char *smask,
                           float *sh_out_x, float *sh_out_y)
{
	int const sx = (sp->lpos[0] < 0.0) ? -1 : 1, sy = (sp->lpos[1] < 0.0) ? -1 : 1;
	float *ox = (float *)malloc((size_t)ntiles*zvsize*sizeof(float)), *oy = (float *)malloc((size_t)ntiles*zvsize*sizeof(float));
	int *done = (int *)calloc(ntiles, sizeof(int));
	for (size_t i = 0; i < (size_t)ntiles*zvsize; ++i) {ox[i] = TW_MESH_MIN_Z; oy[i] = TW_MESH_MIN_Z;} /* sh_out[l][!d].resize(zvsize, MESH_MIN_Z), :677 */
	for (unsigned pass = 0, ndone = 0; ndone < ntiles && pass <= ntiles; ++pass) {
		for (unsigned t = 0; t < ntiles; ++t) {
			if (done[t]) continue;
			int nbx = -1, nby = -1, wait = 0;
			for (unsigned u = 0; u < ntiles; ++u) {
				if (tile_xy[2*u] == tile_xy[2*t] + sx && tile_xy[2*u+1] == tile_xy[2*t+1]) {nbx = (int)u;}
				if (tile_xy[2*u] == tile_xy[2*t] && tile_xy[2*u+1] == tile_xy[2*t+1] + sy) {nby = (int)u;}
			}
			if ((nbx >= 0 && !done[nbx]) || (nby >= 0 && !done[nby])) {wait = 1;}
			if (wait) continue;
			to_calc_mesh_shadows(sp, zvals + (size_t)t*zvsize*zvsize, smask + (size_t)t*zvsize*zvsize, (int)zvsize, (int)zvsize,
			                     (nby >= 0) ? ox + (size_t)nby*zvsize : NULL, (nbx >= 0) ? oy + (size_t)nbx*zvsize : NULL, ox + (size_t)t*zvsize, oy + (size_t)t*zvsize);
			done[t] = 1; ++ndone;
		}
	}
	if (sh_out_x) {memcpy(sh_out_x, ox, (size_t)ntiles*zvsize*sizeof(float));}
	if (sh_out_y) {memcpy(sh_out_y, oy, (size_t)ntiles*zvsize*sizeof(float));}
	free(ox); free(oy); free(done);
}

/* ------------------------------------------------------------------------------------------------ terrain weights texture
 * tile_t::create_texture, src/tiled_mesh.cpp:1071-1248, for terrain without cities / tunnels / buildings (check_mesh_mask == check_city == 0, no exclude cubes,
 * check_buildings == 0), restated loop for loop. Unsuffixed literals are doubles in the reference and here. Pinned against the function itself, cut out of the
 * reference at build time (tests/test_oracle_vs_reference.py::test_tile_weights_against_extracted_create_texture). */
static void tw_update_lttex_ix(int *ix, const tw_weight_params *W) { /* src/Textures.cpp:1289-1292 */
	if (W->snow_to_rock && W->tex_class[*ix] == TW_TEX_SNOW) {--*ix;}
	if (W->vegetation == 0.0 && W->tex_class[*ix] == TW_TEX_GROUND) {++*ix;}
}
static void tw_get_tids(float relh, int *k1, int *k2, float *t, const tw_weight_params *W) { /* src/Textures.cpp:1294-1312 */
	float const TEXTURE_SMOOTH = 0.01; /* :12 */
	const float *h_dirt = W->h_dirt;
	if      (relh < h_dirt[0]) {*k1 = 0;}
	else if (relh < h_dirt[1]) {*k1 = 1;}
	else if (relh < h_dirt[2]) {*k1 = 2;}
	else if (relh < h_dirt[3]) {*k1 = 3;}
	else                       {*k1 = 4;}
	if (*k1 < 5-1 && (h_dirt[*k1] - relh) < TEXTURE_SMOOTH) {
		if (t) {*t = 1.0 - (h_dirt[*k1] - relh)/TEXTURE_SMOOTH;}
		*k2 = *k1+1;
		tw_update_lttex_ix(k1, W);
		tw_update_lttex_ix(k2, W);
	}
	else {
		tw_update_lttex_ix(k1, W);
		*k2 = *k1;
	}
}
#define TW_CLIP_TO_01(x) std_max(0.0f, std_min(1.0f, (x)))
#define TW_BILINEAR(c, x, y) ((y)*((x)*(c)[3] + (1.0f-(x))*(c)[2]) + (1.0f-(y))*((x)*(c)[1] + (1.0f-(x))*(c)[0])) /* BILINEAR_INTERP, :189; c = [y][x] */
void to_tile_weights(const float *zvals_all, const float *rand_all, unsigned ntiles, unsigned zvsize, const float *tile_params, const tw_weight_params *W, unsigned char *rgba,
                     unsigned char *has_any_grass_out)
{
	unsigned const tsize = zvsize - 1;
	int sand_tex_ix = -1, dirt_tex_ix = -1, grass_tex_ix = -1, rock_tex_ix = -1;
	for (int i = 0; i < 5; ++i) { /* get_texture_ixs, :1049-1062 */
		if (W->tex_class[i] == TW_TEX_SAND) sand_tex_ix = i;
		if (W->tex_class[i] == TW_TEX_DIRT) dirt_tex_ix = i;
		if (W->tex_class[i] == TW_TEX_GROUND) grass_tex_ix = i;
		if (W->tex_class[i] == TW_TEX_ROCK) rock_tex_ix = i;
	}
	float const (*sthresh)[2] = W->sthresh;
	float const zmin = W->zmin, zmax = W->zmax, relh_adj_tex = W->relh_adj_tex, water_level = W->water_level, vegetation = W->vegetation;
	float const xy_mult = W->xy_mult;
	float const dz_inv = 1.0f/(zmax - zmin);
	float const steep_mult_grass = 1.0f/(sthresh[0][1] - sthresh[0][0]);
	float const steep_mult_snow  = 1.0f/(sthresh[1][1] - sthresh[1][0]);
	float const steep_mult_rock  = 1.0f/(0.8f*sthresh[0][0] - 0.5f*sthresh[0][0]);
	float const vnz_scale = W->vnz_scale;
	for (unsigned tile = 0; tile < ntiles; ++tile) {
		const float *zvals = zvals_all + (size_t)tile*zvsize*zvsize, *rand_vals = rand_all + (size_t)tile*tsize*tsize, *prm = tile_params + (size_t)tile*8;
		unsigned char *mesh_weight_data = rgba + (size_t)tile*tsize*tsize*4;
		int has_any_grass = 0;
		for (unsigned y = 0; y < tsize; ++y) {
			float const yv = (float)y*xy_mult;
			for (unsigned x = 0; x < tsize; ++x) {
				unsigned const ix_val = y*tsize + x, off = 4*ix_val, ix = y*zvsize + x;
				float weights[5] = {0, 0, 0, 0, 0};
				float const mh00 = zvals[ix], mh01 = zvals[ix+1], mh10 = zvals[ix+zvsize], mh11 = zvals[ix+zvsize+1];
				float const mhmin = std_min(std_min(mh00, mh01), std_min(mh10, mh11)), mhmax = std_max(std_max(mh00, mh01), std_max(mh10, mh11));
				float const rand_offset = W->noise_scale*rand_vals[y*tsize + x]; /* rand_vals[] = noise_scale*height_gen.eval_index(x, y, 50), :1120 */
				float const relh1 = relh_adj_tex + (mhmin - zmin)*dz_inv + rand_offset, relh2 = relh_adj_tex + (mhmax - zmin)*dz_inv + rand_offset;
				int k1, k2, k3, k4;
				tw_get_tids(relh1, &k1, &k2, NULL, W);
				tw_get_tids(relh2, &k3, &k4, NULL, W);
				int const same_tid = (k1 == k4);
				float t = 0.0;
				k2 = k4;
				if (!same_tid) {
					float const relh = relh_adj_tex + (mh00 - zmin)*dz_inv;
					tw_get_tids(relh, &k1, &k2, &t, W);
				}
				float weight_scale = 1.0;
				int const grass = (W->tex_class[k1] == TW_TEX_GROUND || W->tex_class[k2] == TW_TEX_GROUND), snow = (W->tex_class[k2] == TW_TEX_SNOW);
				has_any_grass |= grass;
				if (grass || snow) {
					float const *const sti = sthresh[snow];
					float const nx = W->dy_val*(zvals[ix] - zvals[ix + 1]), ny = W->dx_val*(zvals[ix] - zvals[ix + zvsize]), nz = W->dxdy; /* get_norm_not_normalized, src/tiled_mesh.h:281 */
					float vnz = vnz_scale*nz/sqrtf(nx*nx + ny*ny + nz*nz);
					if (grass && vnz > sti[1]) {vnz = TW_CLIP_TO_01(1.0f + 20.0f*rand_offset);}
					if (vnz < sti[1]) {
						if (grass) {
							float rock_weight = (W->tex_class[k1] == TW_TEX_GROUND || W->tex_class[k2] == TW_TEX_ROCK) ? t : 0.0;
							float const steepness = 1.0 - TW_CLIP_TO_01((vnz - 0.5f*sti[0])*steep_mult_rock);
							rock_weight  = rock_weight*(1.0 - steepness) + steepness;
							weight_scale = TW_CLIP_TO_01((vnz - sti[0])*steep_mult_grass);
							weights[rock_tex_ix] += (1.0 - weight_scale)*rock_weight;
							weights[dirt_tex_ix] += (1.0 - weight_scale)*(1.0 - rock_weight);
						}
						else {
							weight_scale = TW_CLIP_TO_01(2.0f*(vnz - sti[0])*steep_mult_snow);
							weights[rock_tex_ix] += 1.0 - weight_scale;
						}
					}
				}
				weights[k2] += weight_scale*t;
				weights[k1] += weight_scale*(1.0 - t);
				float const xv = (float)x*xy_mult;
				if (vegetation > 0.0) {
					float const dirt_scale = TW_BILINEAR(prm + 4, xv, yv);
					if (dirt_scale < 1.0) {
						weights[sand_tex_ix] += (1.0 - dirt_scale)*weights[dirt_tex_ix];
						weights[dirt_tex_ix] *= dirt_scale;
					}
				}
				if (grass) {
					float grass_scale = (mhmin < water_level) ? 0.0f : TW_BILINEAR(prm, xv, yv);
					if (grass_scale < 1.0) {
						float const gscale = TW_CLIP_TO_01(2.5f*(grass_scale - 0.5f) + 0.5f);
						weights[sand_tex_ix ] += (1.0 - gscale)*weights[grass_tex_ix];
						weights[grass_tex_ix] *= gscale;
					}
				}
				for (unsigned i = 0; i < 5-1; ++i) {
					mesh_weight_data[off+i] = ((weights[i] <= 0.01) ? 0 : ((weights[i] >= 0.99) ? 255 : (unsigned char)(255.0*weights[i])));
				}
			}
		}
		if (has_any_grass_out) {has_any_grass_out[tile] = (unsigned char)has_any_grass;}
	}
}

/* ------------------------------------------------------------------------------------------------ M_SPEC protocol model (test infrastructure)
 * An executable statement of the PROTOCOL behind twi_erode_spec (csrc/tw_erosion.cu): the reference's serial droplet order, walked speculatively
 * and committed in order. Sequential C, nothing shared with the CUDA code: a window of B droplets, every walker reads the committed map through a private overlay of
 * its own writes and records the conflict tiles it touches; per round the walkers run in a PSEUDO-RANDOM order for at most `cap` moves each (suspension), the head of
 * the window walks directly on the map; then stamp (lowest toucher per tile) / validate / commit the prefix up to the first conflict / re-walk what a committed droplet
 * touched. tests/test_oracle_golden.py::test_spec_protocol_model checks that, whatever the window, cap, tile size and schedule, the result is to_apply_erosion's, bit for
 * bit - on maps small enough that almost every droplet conflicts. The per-move arithmetic is to_apply_erosion's (src/erosion.cpp:76-152). */
#ifndef CLAMPI
#define CLAMPI(v, hi) (((v) < (hi)) ? (((v) > 0) ? (v) : 0) : (hi))
#endif
typedef struct {
	unsigned it; int status;                 /* 0 empty, 1 dirty (to be walked from the start), 2 valid (finished, uncommitted), 5 walking (suspended) */
	int inplace;                             /* walked as the head: directly on the map */
	int started, xi, zi; unsigned numMoves;
	float xp, zp, xf, zf, s, v, w, dx, dz, h, h00, h10, h01, h11;
	tw_rng rgen;
	size_t nov; unsigned *ov_cell; float *ov_val; int *ov_pos;   /* overlay of the walker's own writes: list + per-cell position (-1: none) */
	size_t ntile, cap_tile; unsigned *tiles;
	unsigned long long steps;
} spm_walker;
typedef struct {float *mh; int NX, NY, xsize, ysize, shift, TNX; const tw_erosion_params *ep; unsigned MAX_PATH_LEN;} spm_ctx;

static float spm_read(const spm_ctx *C, spm_walker *W, int x, int y) {
	size_t const ix = (size_t)C->NX*CLAMPI(y, C->NY-1) + CLAMPI(x, C->NX-1);
	if (!W->inplace && W->ov_pos[ix] >= 0) return W->ov_val[W->ov_pos[ix]];
	return C->mh[ix];
}
static void spm_add(const spm_ctx *C, spm_walker *W, size_t ix, float delta) { /* heightmap[ix] += delta, privately unless this is the head */
	if (W->inplace) {C->mh[ix] += delta; return;}
	if (W->ov_pos[ix] < 0) {W->ov_pos[ix] = (int)W->nov; W->ov_cell[W->nov] = (unsigned)ix; W->ov_val[W->nov] = C->mh[ix]; ++W->nov;}
	W->ov_val[W->ov_pos[ix]] += delta;
}
static void spm_touch(const spm_ctx *C, spm_walker *W, int xi, int zi) { /* the tiles of [xi-1, xi+2] x [zi-1, zi+2]: every cell this move can read or write */
	int const qx = CLAMPI(xi, C->NX-1), qz = CLAMPI(zi, C->NY-1);
	int const ax = CLAMPI(qx-1, C->NX-1) >> C->shift, bx = CLAMPI(qx+2, C->NX-1) >> C->shift, az = CLAMPI(qz-1, C->NY-1) >> C->shift, bz = CLAMPI(qz+2, C->NY-1) >> C->shift;
	for (int tz = az; tz <= bz; ++tz) for (int tx = ax; tx <= bx; ++tx) {
		unsigned const t = (unsigned)(tz*C->TNX + tx);
		int dup = 0;
		for (size_t k = (W->ntile > 8) ? W->ntile - 8 : 0; k < W->ntile; ++k) {if (W->tiles[k] == t) {dup = 1;}}
		if (dup) continue;
		if (W->ntile == W->cap_tile) {W->cap_tile *= 2; W->tiles = (unsigned *)realloc(W->tiles, W->cap_tile*sizeof(unsigned));}
		W->tiles[W->ntile++] = t;
	}
}
static void spm_reset(const spm_ctx *C, spm_walker *W) { /* forget a walk */
	for (size_t k = 0; k < W->nov; ++k) {W->ov_pos[W->ov_cell[k]] = -1;}
	W->nov = 0; W->ntile = 0; W->started = 0; W->steps = 0; W->inplace = 0;
	(void)C;
}
/* walks at most `cap` moves; returns 1 when the droplet has ended */
static int spm_walk(const spm_ctx *C, spm_walker *W, unsigned cap) {
	float const Kq=10, Kw=0.001f, Kr=0.9f, Kd=0.02f, Ki=0.1f, minSlope=0.05f, g=20, Kg=g*2;
	int const PAD = 4, NX = C->NX, NY = C->NY;
	const tw_erosion_params *ep = C->ep;
	float const erode_amount = ep->erode_amount, tp = two_pi();
	if (!W->started) {
		to_rng_set(&W->rgen, (int)W->it+11, 79*(int)W->it+121);
		W->xi = PAD + (to_rng_rand(&W->rgen)%C->xsize);
		W->zi = PAD + (to_rng_rand(&W->rgen)%C->ysize);
		W->xp=W->xi; W->zp=W->zi; W->xf=0; W->zf=0; W->s=0; W->v=0; W->w=1; W->dx=0; W->dz=0;
		spm_touch(C, W, W->xi, W->zi);
		W->h=spm_read(C, W, W->xi, W->zi); W->h00=W->h; W->h10=spm_read(C, W, W->xi+1, W->zi); W->h01=spm_read(C, W, W->xi, W->zi+1); W->h11=spm_read(C, W, W->xi+1, W->zi+1);
		W->numMoves = 0; W->started = 1;
	}
	int xi=W->xi, zi=W->zi; float xp=W->xp, zp=W->zp, xf=W->xf, zf=W->zf, s=W->s, v=W->v, w=W->w, dx=W->dx, dz=W->dz, h=W->h, h00=W->h00, h10=W->h10, h01=W->h01, h11=W->h11;
	int ended = 0;
	unsigned moved = 0;
#define SPM_DEPOSIT_AT(X, Z, WGT) {float const delta = ds*erode_amount*(WGT); if (!((X) < 0 || (Z) < 0 || (X) >= NX || (Z) >= NY)) {spm_add(C, W, (size_t)NX*CLAMPI((Z), NY-1) + CLAMPI((X), NX-1), delta);}}
#define SPM_DEPOSIT(H) SPM_DEPOSIT_AT(xi, zi, (1-xf)*(1-zf)) SPM_DEPOSIT_AT(xi+1, zi, xf*(1-zf)) SPM_DEPOSIT_AT(xi, zi+1, (1-xf)*zf) SPM_DEPOSIT_AT(xi+1, zi+1, xf*zf) (H)+=ds;
	for (;;) {
		if (W->numMoves >= C->MAX_PATH_LEN) {ended = 1; break;}
		if (moved >= cap) break; /* suspended */
		++W->numMoves; ++moved; ++W->steps;
		spm_touch(C, W, xi, zi);
		float gx=h00+h01-h10-h11, gz=h00+h10-h01-h11;
		dx=(dx-gx)*Ki+gx;
		dz=(dz-gz)*Ki+gz;
		float dl=sqrtf(dx*dx+dz*dz);
		if (dl<=FLT_EPSILON) {float a=to_rng_rand_float(&W->rgen)*tp; dx=cosf(a); dz=sinf(a);}
		else {dx/=dl; dz/=dl;}
		float nxp=xp+dx, nzp=zp+dz;
		int nxi, nzi;
		if (fabsf(nxp) < 2147483648.0f && fabsf(nzp) < 2147483648.0f) {nxi=(int)floorf(nxp); nzi=(int)floorf(nzp);}
		else {nxi = nzi = (-2147483647 - 1); if (fabsf(nxp) < 2147483648.0f) {nxi=(int)floorf(nxp);} if (fabsf(nzp) < 2147483648.0f) {nzi=(int)floorf(nzp);} /* x86 cvttss2si */
			if (W->ntile == W->cap_tile) {W->cap_tile *= 2; W->tiles = (unsigned *)realloc(W->tiles, W->cap_tile*sizeof(unsigned));}
			W->tiles[W->ntile++] = 0u; /* the reference reads cell (0, 0) next */
		}
		float nxf=nxp-nxi, nzf=nzp-nzi;
		float nh00=spm_read(C, W, nxi, nzi), nh10=spm_read(C, W, (nxi == (-2147483647 - 1)) ? nxi : nxi+1, nzi), nh01=spm_read(C, W, nxi, (nzi == (-2147483647 - 1)) ? nzi : nzi+1),
		      nh11=spm_read(C, W, (nxi == (-2147483647 - 1)) ? nxi : nxi+1, (nzi == (-2147483647 - 1)) ? nzi : nzi+1);
		float nh=(nh00*(1-nxf)+nh10*nxf)*(1-nzf)+(nh01*(1-nxf)+nh11*nxf)*nzf;
		if (std_max(std_max(nh00, nh10), std_max(nh01, nh11)) < ep->water_plane_z - ep->half_dxy) {ended = 1; break;}
		int const outside = (xi < 0 || zi < 0 || xi >= NX || zi >= NY);
		if (nh>=h || outside) {
			float ds=(nh-h)+0.001f;
			if (ds>=s || outside) {ds=s; SPM_DEPOSIT(h) s=0; ended = 1; break;}
			SPM_DEPOSIT(h)
			s-=ds;
			v=0;
		}
		float dh=h-nh;
		float q=std_max(dh, minSlope)*v*w*Kq;
		float ds=s-q;
		if (ds>=0) {ds*=Kd; SPM_DEPOSIT(dh) s-=ds;}
		else {
			ds*=-Kr;
			ds=std_min(ds, dh*0.99f);
			{float const relh = ep->relh_adj_tex + (nh - ep->zmin)/(ep->zmax - ep->zmin); ds = (float)(ds*((relh > ep->clip_hd1) ? 0.5 : 2.0));}
			for (int z=zi-1; z<=zi+2; ++z) {
				float zo=z-zp, zo2=zo*zo;
				for (int x=xi-1; x<=xi+2; ++x) {
					float xo=x-xp;
					float wgt=1-(xo*xo+zo2)*0.25f;
					if (wgt<=0) continue;
					wgt*=0.1591549430918953f;
					float const delta=ds*erode_amount*wgt;
					spm_add(C, W, (size_t)NX*CLAMPI(z, NY-1) + CLAMPI(x, NX-1), -delta);
				}
			}
			dh-=ds;
			s+=ds;
		}
		v=sqrtf(v*v+Kg*dh);
		w*=1-Kw;
		xp=nxp; zp=nzp; xi=nxi; zi=nzi; xf=nxf; zf=nzf;
		h=nh; h00=nh00; h10=nh10; h01=nh01; h11=nh11;
	}
	W->xi=xi; W->zi=zi; W->xp=xp; W->zp=zp; W->xf=xf; W->zf=zf; W->s=s; W->v=v; W->w=w; W->dx=dx; W->dz=dz; W->h=h; W->h00=h00; W->h10=h10; W->h01=h01; W->h11=h11;
	return ended;
}
/* returns the droplet moves (of the walks that were committed); stats[0..3] (optional) = rounds, walks started, walks thrown away, droplets walked in place */
unsigned long long to_erode_spec_model(float *heightmap, int xsize, int ysize, float min_zval, unsigned num_iters, const tw_erosion_params *ep,
                                       unsigned B, unsigned cap, int tile_shift, unsigned sched_seed, unsigned long long *stats)
{
	if (num_iters == 0 || ep->erode_amount <= 0.0 || B == 0 || cap == 0) return 0;
	int const PAD = 4, NX = xsize+2*PAD, NY = ysize+2*PAD;
	spm_ctx C;
	C.NX = NX; C.NY = NY; C.xsize = xsize; C.ysize = ysize; C.shift = tile_shift; C.TNX = ((NX-1) >> tile_shift) + 1; C.ep = ep; C.MAX_PATH_LEN = 4u*(unsigned)NX*(unsigned)NY;
	C.mh = (float *)malloc((size_t)NX*NY*sizeof(float));
	for (int y = 0; y < NY; ++y) {for (int x = 0; x < NX; ++x) {C.mh[(size_t)y*NX + x] = heightmap[CLAMPI(x-PAD, xsize-1) + (size_t)CLAMPI(y-PAD, ysize-1)*xsize];}}
	size_t const ntile = (size_t)C.TNX*(((NY-1) >> tile_shift) + 1);
	unsigned *stamps = (unsigned *)malloc(ntile*sizeof(unsigned));
	for (size_t t = 0; t < ntile; ++t) {stamps[t] = 0xffffffffu;}
	spm_walker *Wk = (spm_walker *)calloc(B, sizeof(spm_walker));
	for (unsigned s = 0; s < B; ++s) {
		spm_walker *W = &Wk[s];
		W->it = s; W->status = (s < num_iters) ? 1 : 0;
		W->ov_cell = (unsigned *)malloc((size_t)NX*NY*sizeof(unsigned)); W->ov_val = (float *)malloc((size_t)NX*NY*sizeof(float)); W->ov_pos = (int *)malloc((size_t)NX*NY*sizeof(int));
		for (size_t i = 0; i < (size_t)NX*NY; ++i) {W->ov_pos[i] = -1;}
		W->cap_tile = 64; W->tiles = (unsigned *)malloc(W->cap_tile*sizeof(unsigned));
	}
	unsigned long long steps = 0, rounds = 0, walks = 0, wasted = 0, inplace = 0;
	unsigned lo = 0, rnd = sched_seed*2654435761u + 12345u;
	unsigned *order = (unsigned *)malloc(B*sizeof(unsigned));
	while (lo < num_iters) {
		++rounds;
		unsigned const hi = (num_iters - lo < B) ? num_iters : lo + B;
		/* ---- walk, in a pseudo-random order of the slots (the GPU's walkers run concurrently: the head's in-place writes land before, between or after the others' reads) */
		for (unsigned s = 0; s < B; ++s) {order[s] = s;}
		for (unsigned s = B; s > 1; --s) {rnd = rnd*1664525u + 1013904223u; unsigned const j = (rnd >> 8) % s, t = order[s-1]; order[s-1] = order[j]; order[j] = t;}
		for (unsigned k = 0; k < B; ++k) {
			spm_walker *W = &Wk[order[k]];
			if (W->it >= num_iters || !(W->status == 1 || W->status == 5)) continue;
			if (W->status == 1) {spm_reset(&C, W); ++walks; W->inplace = (W->it == lo); if (W->inplace) {++inplace;}}
			W->status = spm_walk(&C, W, W->inplace ? 2*cap : cap) ? 2 : 5;
		}
		/* ---- stamp */
		for (unsigned s = 0; s < B; ++s) {
			spm_walker *W = &Wk[s];
			if (W->it >= num_iters || !(W->status == 2 || W->status == 5)) continue;
			for (size_t k = 0; k < W->ntile; ++k) {if (W->it < stamps[W->tiles[k]]) {stamps[W->tiles[k]] = W->it;}}
		}
		/* ---- validate: the first droplet that cannot be committed */
		unsigned f = hi;
		unsigned *minw = order; /* reuse */
		for (unsigned s = 0; s < B; ++s) {
			spm_walker *W = &Wk[s];
			minw[s] = 0xffffffffu;
			if (W->it >= num_iters || W->status == 0) continue;
			if (!(W->status == 2 || W->status == 5)) {if (W->it < f) {f = W->it;} continue;}
			for (size_t k = 0; k < W->ntile; ++k) {if (stamps[W->tiles[k]] < minw[s]) {minw[s] = stamps[W->tiles[k]];}}
			if ((minw[s] < W->it || W->status == 5) && W->it < f) {f = W->it;}
		}
		/* ---- commit the prefix [lo, f), re-walk what a committed droplet touched, take the stamps back */
		for (unsigned s = 0; s < B; ++s) {
			spm_walker *W = &Wk[s];
			if (W->it >= num_iters || W->status == 0) continue;
			if (W->status == 2 || W->status == 5) {for (size_t k = 0; k < W->ntile; ++k) {stamps[W->tiles[k]] = 0xffffffffu;}}
			if (W->it < f) {
				if (!W->inplace) {for (size_t k = 0; k < W->nov; ++k) {C.mh[W->ov_cell[k]] = W->ov_val[k];}}
				steps += W->steps;
				spm_reset(&C, W);
				W->it += B; W->status = (W->it < num_iters) ? 1 : 0;
			}
			else if ((W->status == 2 || W->status == 5) && minw[s] < f) {W->status = 1; ++wasted;}
		}
		if (rounds > 64ull*num_iters*(1ull + C.MAX_PATH_LEN/cap)) {free(order); order = NULL; break;} /* no progress (cannot happen: a finished head is never in conflict, and the head finishes) */
		lo = f;
	}
	int const ok = (order != NULL);
	for (int y = 0; y < ysize; ++y) {for (int x = 0; x < xsize; ++x) {heightmap[(size_t)y*xsize + x] = std_max(min_zval, C.mh[(size_t)(y+PAD)*NX + x+PAD]);}}
	for (unsigned s = 0; s < B; ++s) {free(Wk[s].ov_cell); free(Wk[s].ov_val); free(Wk[s].ov_pos); free(Wk[s].tiles);}
	free(Wk); free(stamps); free(order); free(C.mh);
	if (stats) {stats[0] = rounds; stats[1] = walks; stats[2] = wasted; stats[3] = inplace;}
	return ok ? steps : ~0ull;
}

