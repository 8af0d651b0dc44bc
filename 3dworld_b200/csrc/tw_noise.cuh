// tw_noise.cuh - device restatement of the reference's gradient noise (GLM 0.9.9.1 gtc/noise, Ashima/Gustavson arithmetic hash):
//   glm::simplex(vec2)  dependencies/glm/glm/gtc/noise.inl:592-646     glm::perlin(vec2)  :25-62
//   glm::simplex(vec3)  :649-721                                        glm::perlin(vec3)  :66-133
//   helpers             dependencies/glm/glm/detail/_noise.hpp:15-84, detail/func_common.inl:19,28,86,188,216,252
// The reference is built without FMA contraction (makefile:11, no -march), so every a*b+c below is a separate multiply and add
// (this TU is compiled with -fmad=false). __fmaf_rn is used ONLY where the product is exactly representable (small integers, or a
// power-of-two factor), in which case the fused and unfused results are identical bit for bit:
//   * permute(): x in [0,289) integer-valued, so x*34+1 <= 9827 and (x*34+1)*x < 2^24 are exact;
//   * mod289():  floor(..)*289 is an integer < 2^24, so x - t*289 has one rounding either way (and is itself exact);
//   * 2*f - 1:   doubling is exact.
#pragma once
#include <cuda_runtime.h>

namespace twn {

__device__ __forceinline__ float mod289(float x)  {float const t = floorf(x*(1.0f/289.0f)); return __fmaf_rn(-t, 289.0f, x);}
__device__ __forceinline__ float permute(float x) {return mod289(__fmaf_rn(x, 34.0f, 1.0f)*x);}
// glm::mod(a, 289) = a - 289*floor(a/289) with a true division (func_common.inl:216), a integer-valued (a floor() result).
// mod_div289: the literal form. mod_int289: division-free form that is bit-identical for every integer |a| < 2^22 (verified exhaustively
// over all 8.4 M integers, tests/test_host_logic.py::test_division_free_forms): q = floor(a*RN(1/289)) is either floor(a/289) or, only
// when a is a multiple of 289, one less (RN(1/289) > 1/289 and the product error stays below 1/289), so r = a - 289q (exact) needs one
// conditional subtract. Callers guard the range and fall back to mod_div289 for astronomically large lattice indices.
__device__ __forceinline__ float mod_div289(float a) {float const t = floorf(__fdiv_rn(a, 289.0f)); return a - 289.0f*t;}
__device__ __forceinline__ float mod_int289(float a) {
	float const q = floorf(a*(1.0f/289.0f));
	float const r = __fmaf_rn(q, -289.0f, a); // exact: |q*289| < 2^23
	return (r >= 289.0f) ? r - 289.0f : r;
}
constexpr float MOD_FAST_LIMIT = 4194304.0f; // 2^22
// i/41 for integer i in [0,289): quotient by one Newton correction of i*RN(1/41); equals the IEEE quotient for all 289 inputs (same test)
__device__ __forceinline__ float div41_small(float i) {
	float const c = 1.0f/41.0f, q0 = i*c;
	return __fmaf_rn(__fmaf_rn(-q0, 41.0f, i), c, q0);
}
__device__ __forceinline__ float tinvsqrt(float r) {return 1.79284291400159f - 0.85373472095314f*r;}
__device__ __forceinline__ float fade(float t)    {return (t*t*t)*(t*(t*6.0f - 15.0f) + 10.0f);}
__device__ __forceinline__ float fract(float x)   {return x - floorf(x);}
__device__ __forceinline__ float mix(float x, float y, float a) {return x + a*(y - x);}
__device__ __forceinline__ float gmax(float x, float y) {return (x < y) ? y : x;}   // glm::max
// glm::max(x, 0) for the simplex falloff terms: fmaxf differs from (x < 0) ? 0 : x only when x is NaN, and then the final product is NaN either way
__device__ __forceinline__ float gmax0(float x) {return fmaxf(x, 0.0f);}
__device__ __forceinline__ float gmin(float x, float y) {return (y < x) ? y : x;}   // glm::min
__device__ __forceinline__ float step(float edge, float x) {return (x < edge) ? 0.0f : 1.0f;}
__device__ __forceinline__ float two_f_minus_1(float f) {return __fmaf_rn(2.0f, f, -1.0f);}

__device__ __forceinline__ float simplex2(float vx, float vy) {
	float const Cx = 0.211324865405187f, Cy = 0.366025403784439f, Cz = -0.577350269189626f, Cw = 0.024390243902439f;
	float const s = vx*Cy + vy*Cy;
	float ix = floorf(vx + s), iy = floorf(vy + s);
	float const t = ix*Cx + iy*Cx;
	float const x0x = vx - ix + t, x0y = vy - iy + t;
	bool  const xgt = (x0x > x0y);
	float const i1x = xgt ? 1.0f : 0.0f, i1y = xgt ? 0.0f : 1.0f;
	float const x12x = (x0x + Cx) - i1x, x12y = (x0y + Cx) - i1y, x12z = x0x + Cz, x12w = x0y + Cz;
	// results are in [0,289) and never -0, so the reference's "+ 0.0f" terms are no-ops
	if (fmaxf(fabsf(ix), fabsf(iy)) < MOD_FAST_LIMIT) {ix = mod_int289(ix); iy = mod_int289(iy);} else {ix = mod_div289(ix); iy = mod_div289(iy);}
	float const q0 = permute(iy), q1 = permute(iy + i1y), q2 = permute(iy + 1.0f);
	float const p0 = permute(q0 + ix), p1 = permute(q1 + ix + i1x), p2 = permute(q2 + ix + 1.0f);
	float m0 = gmax0(0.5f - (x0x*x0x + x0y*x0y));
	float m1 = gmax0(0.5f - (x12x*x12x + x12y*x12y));
	float m2 = gmax0(0.5f - (x12z*x12z + x12w*x12w));
	m0 = m0*m0; m1 = m1*m1; m2 = m2*m2;
	m0 = m0*m0; m1 = m1*m1; m2 = m2*m2;
	float const X0 = two_f_minus_1(fract(p0*Cw)), X1 = two_f_minus_1(fract(p1*Cw)), X2 = two_f_minus_1(fract(p2*Cw));
	float const h0 = fabsf(X0) - 0.5f, h1 = fabsf(X1) - 0.5f, h2 = fabsf(X2) - 0.5f;
	float const a0 = X0 - floorf(X0 + 0.5f), a1 = X1 - floorf(X1 + 0.5f), a2 = X2 - floorf(X2 + 0.5f);
	m0 *= 1.79284291400159f - 0.85373472095314f*(a0*a0 + h0*h0);
	m1 *= 1.79284291400159f - 0.85373472095314f*(a1*a1 + h1*h1);
	m2 *= 1.79284291400159f - 0.85373472095314f*(a2*a2 + h2*h2);
	float const gx = a0*x0x + h0*x0y, gy = a1*x12x + h1*x12y, gz = a2*x12z + h2*x12w;
	return 130.0f*(m0*gx + m1*gy + m2*gz);
}

__device__ __forceinline__ void perlin2_corner(float ix, float iy, float &gx, float &gy) {
	float const i = permute(permute(ix) + iy);
	float const g = two_f_minus_1(fract(div41_small(i)));
	gy = fabsf(g) - 0.5f;
	gx = g - floorf(g + 0.5f);
}

__device__ __forceinline__ float perlin2(float Px, float Py) {
	float const flx = floorf(Px), fly = floorf(Py);
	float const frx = Px - flx, fry = Py - fly;           // fract()
	float const Pfz = frx - 1.0f, Pfw = fry - 1.0f;
	float Pix, Piy, Piz, Piw;
	if (fmaxf(fabsf(flx), fabsf(fly)) < MOD_FAST_LIMIT - 1.0f) {Pix = mod_int289(flx); Piy = mod_int289(fly); Piz = mod_int289(flx + 1.0f); Piw = mod_int289(fly + 1.0f);}
	else {Pix = mod_div289(flx); Piy = mod_div289(fly); Piz = mod_div289(flx + 1.0f); Piw = mod_div289(fly + 1.0f);}
	float g00x, g00y, g10x, g10y, g01x, g01y, g11x, g11y;
	perlin2_corner(Pix, Piy, g00x, g00y);
	perlin2_corner(Piz, Piy, g10x, g10y);
	perlin2_corner(Pix, Piw, g01x, g01y);
	perlin2_corner(Piz, Piw, g11x, g11y);
	float const n_00 = tinvsqrt(g00x*g00x + g00y*g00y), n_01 = tinvsqrt(g01x*g01x + g01y*g01y);
	float const n_10 = tinvsqrt(g10x*g10x + g10y*g10y), n_11 = tinvsqrt(g11x*g11x + g11y*g11y);
	g00x *= n_00; g00y *= n_00; g01x *= n_01; g01y *= n_01; g10x *= n_10; g10y *= n_10; g11x *= n_11; g11y *= n_11;
	float const n00 = g00x*frx + g00y*fry;
	float const n10 = g10x*Pfz + g10y*fry;
	float const n01 = g01x*frx + g01y*Pfw;
	float const n11 = g11x*Pfz + g11y*Pfw;
	float const fdx = fade(frx), fdy = fade(fry);
	float const nx0 = mix(n00, n10, fdx), nx1 = mix(n01, n11, fdx);
	return 2.3f*mix(nx0, nx1, fdy);
}

__device__ __forceinline__ void perlin3_grad(float ixyz, float &gx, float &gy, float &gz) {
	float const c17 = (float)(1.0/7.0);
	float g = ixyz*c17;
	gy = fract(floorf(g)*c17) - 0.5f;
	gx = fract(g);
	gz = 0.5f - fabsf(gx) - fabsf(gy);
	float const sz = step(gz, 0.0f);
	gx -= sz*(step(0.0f, gx) - 0.5f);
	gy -= sz*(step(0.0f, gy) - 0.5f);
	float const n = tinvsqrt(gx*gx + gy*gy + gz*gz);
	gx *= n; gy *= n; gz *= n;
}

__device__ __forceinline__ float perlin3(float Px, float Py, float Pz) {
	float const flx = floorf(Px), fly = floorf(Py), flz = floorf(Pz);
	float const Pi0x = mod289(flx), Pi0y = mod289(fly), Pi0z = mod289(flz);
	float const Pi1x = mod289(flx + 1.0f), Pi1y = mod289(fly + 1.0f), Pi1z = mod289(flz + 1.0f);
	float const f0x = Px - flx, f0y = Py - fly, f0z = Pz - flz;
	float const f1x = f0x - 1.0f, f1y = f0y - 1.0f, f1z = f0z - 1.0f;
	float const px0 = permute(Pi0x), px1 = permute(Pi1x);
	float const ixy00 = permute(px0 + Pi0y), ixy10 = permute(px1 + Pi0y), ixy01 = permute(px0 + Pi1y), ixy11 = permute(px1 + Pi1y);
	float gx, gy, gz;
	perlin3_grad(permute(ixy00 + Pi0z), gx, gy, gz); float const n000 = gx*f0x + gy*f0y + gz*f0z;
	perlin3_grad(permute(ixy10 + Pi0z), gx, gy, gz); float const n100 = gx*f1x + gy*f0y + gz*f0z;
	perlin3_grad(permute(ixy01 + Pi0z), gx, gy, gz); float const n010 = gx*f0x + gy*f1y + gz*f0z;
	perlin3_grad(permute(ixy11 + Pi0z), gx, gy, gz); float const n110 = gx*f1x + gy*f1y + gz*f0z;
	perlin3_grad(permute(ixy00 + Pi1z), gx, gy, gz); float const n001 = gx*f0x + gy*f0y + gz*f1z;
	perlin3_grad(permute(ixy10 + Pi1z), gx, gy, gz); float const n101 = gx*f1x + gy*f0y + gz*f1z;
	perlin3_grad(permute(ixy01 + Pi1z), gx, gy, gz); float const n011 = gx*f0x + gy*f1y + gz*f1z;
	perlin3_grad(permute(ixy11 + Pi1z), gx, gy, gz); float const n111 = gx*f1x + gy*f1y + gz*f1z;
	float const fx = fade(f0x), fy = fade(f0y), fz = fade(f0z);
	float const nz0 = mix(n000, n001, fz), nz1 = mix(n100, n101, fz), nz2 = mix(n010, n011, fz), nz3 = mix(n110, n111, fz);
	float const nyz0 = mix(nz0, nz2, fy), nyz1 = mix(nz1, nz3, fy);
	return 2.2f*mix(nyz0, nyz1, fx);
}

// gradient + dot for one simplex3 corner pair handled by the caller; helper computes (x,y,h) lattice terms from the hash value
__device__ __forceinline__ void simplex3_xyh(float p, float &X, float &Y, float &H) {
	float const n_ = 0.142857142857f;
	float const nsx = n_*2.0f - 0.0f, nsy = n_*0.5f - 1.0f, nsz = n_*1.0f - 0.0f;
	float const j  = p - 49.0f*floorf(p*nsz*nsz);
	float const x_ = floorf(j*nsz);
	float const y_ = floorf(j - 7.0f*x_);
	X = x_*nsx + nsy; Y = y_*nsx + nsy;
	H = 1.0f - fabsf(X) - fabsf(Y);
}

__device__ __forceinline__ float simplex3(float vx, float vy, float vz) {
	float const Cx = (float)(1.0/6.0), Cy = (float)(1.0/3.0);
	float const s = vx*Cy + vy*Cy + vz*Cy;
	float i0 = floorf(vx + s), i1_ = floorf(vy + s), i2_ = floorf(vz + s);
	float const t = i0*Cx + i1_*Cx + i2_*Cx;
	float const x0x = vx - i0 + t, x0y = vy - i1_ + t, x0z = vz - i2_ + t;
	float const gx = step(x0y, x0x), gy = step(x0z, x0y), gz = step(x0x, x0z);
	float const lx = 1.0f - gx, ly = 1.0f - gy, lz = 1.0f - gz;
	float const i1x = gmin(gx, lz), i1y = gmin(gy, lx), i1z = gmin(gz, ly);
	float const i2x = gmax(gx, lz), i2y = gmax(gy, lx), i2z = gmax(gz, ly);
	float const x1x = x0x - i1x + Cx, x1y = x0y - i1y + Cx, x1z = x0z - i1z + Cx;
	float const x2x = x0x - i2x + Cy, x2y = x0y - i2y + Cy, x2z = x0z - i2z + Cy;
	float const x3x = x0x - 0.5f, x3y = x0y - 0.5f, x3z = x0z - 0.5f;
	i0 = mod289(i0); i1_ = mod289(i1_); i2_ = mod289(i2_);    // mod289 results are never -0 => "+ 0.0f" terms are no-ops
	float const p0 = permute(permute(permute(i2_)        + i1_)        + i0);
	float const p1 = permute(permute(permute(i2_ + i1z)  + i1_ + i1y)  + i0 + i1x);
	float const p2 = permute(permute(permute(i2_ + i2z)  + i1_ + i2y)  + i0 + i2x);
	float const p3 = permute(permute(permute(i2_ + 1.0f) + i1_ + 1.0f) + i0 + 1.0f);
	float X0, Y0, H0, X1, Y1, H1, X2, Y2, H2, X3, Y3, H3;
	simplex3_xyh(p0, X0, Y0, H0); simplex3_xyh(p1, X1, Y1, H1); simplex3_xyh(p2, X2, Y2, H2); simplex3_xyh(p3, X3, Y3, H3);
	// b0=(X0,X1,Y0,Y1) b1=(X2,X3,Y2,Y3); s = floor(b)*2+1; sh = -step(h,0)
	float const sh0 = -step(H0, 0.0f), sh1 = -step(H1, 0.0f), sh2 = -step(H2, 0.0f), sh3 = -step(H3, 0.0f);
	// a0 = (b0.x + s0.x*sh.x, b0.z + s0.z*sh.x, b0.y + s0.y*sh.y, b0.w + s0.w*sh.y)
	float P0x = X0 + (floorf(X0)*2.0f + 1.0f)*sh0, P0y = Y0 + (floorf(Y0)*2.0f + 1.0f)*sh0, P0z = H0;
	float P1x = X1 + (floorf(X1)*2.0f + 1.0f)*sh1, P1y = Y1 + (floorf(Y1)*2.0f + 1.0f)*sh1, P1z = H1;
	float P2x = X2 + (floorf(X2)*2.0f + 1.0f)*sh2, P2y = Y2 + (floorf(Y2)*2.0f + 1.0f)*sh2, P2z = H2;
	float P3x = X3 + (floorf(X3)*2.0f + 1.0f)*sh3, P3y = Y3 + (floorf(Y3)*2.0f + 1.0f)*sh3, P3z = H3;
	float const n0 = tinvsqrt(P0x*P0x + P0y*P0y + P0z*P0z), n1 = tinvsqrt(P1x*P1x + P1y*P1y + P1z*P1z);
	float const n2 = tinvsqrt(P2x*P2x + P2y*P2y + P2z*P2z), n3 = tinvsqrt(P3x*P3x + P3y*P3y + P3z*P3z);
	P0x *= n0; P0y *= n0; P0z *= n0; P1x *= n1; P1y *= n1; P1z *= n1; P2x *= n2; P2y *= n2; P2z *= n2; P3x *= n3; P3y *= n3; P3z *= n3;
	float m0 = gmax0(0.6f - (x0x*x0x + x0y*x0y + x0z*x0z)), m1 = gmax0(0.6f - (x1x*x1x + x1y*x1y + x1z*x1z));
	float m2 = gmax0(0.6f - (x2x*x2x + x2y*x2y + x2z*x2z)), m3 = gmax0(0.6f - (x3x*x3x + x3y*x3y + x3z*x3z));
	m0 = m0*m0; m1 = m1*m1; m2 = m2*m2; m3 = m3*m3;
	float const d0 = P0x*x0x + P0y*x0y + P0z*x0z, d1 = P1x*x1x + P1y*x1y + P1z*x1z;
	float const d2 = P2x*x2x + P2y*x2y + P2z*x2z, d3 = P3x*x3x + P3y*x3y + P3z*x3z;
	return 42.0f*(((m0*m0)*d0 + (m1*m1)*d1) + ((m2*m2)*d2 + (m3*m3)*d3));
}

} // namespace twn
