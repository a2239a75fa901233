// fm_aniso.cuh -- device math of the anisotropic path (sm_100a): RayMarcher::WPCA
// (src/app/AdvancedRenderer/RayMarcher.cpp:114-254), AnisotropicKernel::W / gradW (src/app/Kernel.cpp:55-107),
// CubicKernel::W (:111-125) and the third-party routine WPCA calls,
// Eigen::SelfAdjointEigenSolver<Matrix3f>::computeDirect (vendor/eigen/include/Eigen/src/Eigenvalues/
// SelfAdjointEigenSolver.h:583-741, closed-form 3x3), including the libm calls inside it (std::atan2 / cos / sin:
// glibc 2.39's algorithms -- fdlibm e_atan2f.c / s_atanf.c in FP32, s_sincosf.h polynomials in FP64 -- so that the
// device computes the same bits as the reference build on the same inputs).
//
// Numeric contract: this header is written in plain C expressions, evaluated left to right with one IEEE rounding
// per operation.  That only holds when the translation unit is compiled WITHOUT multiply-add contraction, so the
// header refuses to compile unless the build passes -fmad=false and says so with -DFM_NO_FMAD (csrc/Makefile does that
// for fm_aniso.cu only; divisions and square roots are IEEE by nvcc's defaults -prec-div=true -prec-sqrt=true).
#pragma once

#ifndef FM_NO_FMAD
#error "fm_aniso.cuh needs a translation unit compiled with -fmad=false -DFM_NO_FMAD"
#endif

#include "fm_common.cuh"

namespace fm
{
namespace aniso
{

// ---- glibc 2.39 atan2f / sinf / cosf on the solver's argument ranges -------------------------------------------
// s_atanf.c (fdlibm)
__device__ __forceinline__ float atanf_fdlibm(float x)
{
	const float atanhi[4] = { 4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f };
	const float atanlo[4] = { 5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f };
	const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
		aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
		aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
	int32_t const hx = __float_as_int(x);
	int32_t const ix = hx & 0x7fffffff;
	int id;
	float hi = 0.0f, lo = 0.0f;
	if (ix >= 0x4c000000)   // |x| >= 2^25
	{
		if (ix > 0x7f800000) return x + x;
		return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
	}
	if (ix < 0x3ee00000)    // |x| < 0.4375
	{
		if (ix < 0x31000000) return x;   // |x| < 2^-29
		id = -1;
	}
	else
	{
		x = fabsf(x);
		if (ix < 0x3f980000)
		{
			if (ix < 0x3f300000) { id = 0; hi = atanhi[0]; lo = atanlo[0]; x = (2.0f * x - 1.0f) / (2.0f + x); }
			else { id = 1; hi = atanhi[1]; lo = atanlo[1]; x = (x - 1.0f) / (x + 1.0f); }
		}
		else
		{
			if (ix < 0x401c0000) { id = 2; hi = atanhi[2]; lo = atanlo[2]; x = (x - 1.5f) / (1.0f + 1.5f * x); }
			else { id = 3; hi = atanhi[3]; lo = atanlo[3]; x = -1.0f / x; }
		}
	}
	float const z = x * x;
	float const w = z * z;
	float const s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
	float const s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
	if (id < 0) return x - x * (s1 + s2);
	float const r = hi - ((x * (s1 + s2) - lo) - x);
	return hx < 0 ? -r : r;
}

// e_atan2f.c (fdlibm)
__device__ __forceinline__ float atan2f_fdlibm(float y, float x)
{
	const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
		pi_lo = -8.7422776573e-08f;
	int32_t const hx = __float_as_int(x), hy = __float_as_int(y);
	int32_t const ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
	if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
	if (hx == 0x3f800000) return atanf_fdlibm(y);
	int32_t const m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
	if (iy == 0)
	{
		if (m < 2) return y;
		return m == 2 ? pi + tiny : -pi - tiny;
	}
	if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
	if (ix == 0x7f800000)
	{
		if (iy == 0x7f800000)
		{
			if (m == 0) return pi_o_4 + tiny;
			if (m == 1) return -pi_o_4 - tiny;
			if (m == 2) return 3.0f * pi_o_4 + tiny;
			return -3.0f * pi_o_4 - tiny;
		}
		if (m == 0) return 0.0f;
		if (m == 1) return -0.0f;
		if (m == 2) return pi + tiny;
		return -pi - tiny;
	}
	if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
	int32_t const k = (iy - ix) >> 23;
	float z;
	if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
	else if (hx < 0 && k < -60) z = 0.0f;
	else z = atanf_fdlibm(fabsf(y / x));
	if (m == 0) return z;
	if (m == 1) return __uint_as_float(__float_as_uint(z) ^ 0x80000000u);
	if (m == 2) return pi - (z - pi_lo);
	return (z - pi_lo) - pi;
}

// sincosf.h / sincosf_data.c: polynomials on [-pi/4, pi/4] in double, reduction by multiples of pi/2
__device__ __forceinline__ float sincos_poly(double x, double x2, int n)
{
	const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
		c4 = 0x1.99343027bf8c3p-16;
	const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
	if ((n & 1) == 0)
	{
		double const x3 = x * x2;
		double const t1 = s2 + x2 * s3;
		double const x7 = x3 * x2;
		double const s = x + x3 * s1;
		return (float)(s + x7 * t1);
	}
	double const x4 = x2 * x2;
	double const t2 = c3 + x2 * c4;
	double const t1 = c1 + x2 * c2;
	double const x6 = x4 * x2;
	double const c = c0 + x2 * t1;
	return (float)(c + x6 * t2);
}

__device__ __forceinline__ uint32_t abstop12(float x) { return (__float_as_uint(x) >> 20) & 0x7ffu; }

// which = 0: sinf(y), 1: cosf(y); exact restatement for 0 <= y < 100 (computeRoots passes theta in [0, pi/3]);
// anything else (only NaN can occur) propagates as NaN
__device__ __forceinline__ float sincosf_glibc(float y, int which)
{
	if (!(y >= 0.0f && y < 100.0f)) return y + y;
	double x = (double)y;
	if (abstop12(y) < abstop12(0x1.921FB6p-1f))
	{
		if (abstop12(y) < abstop12(0x1p-12f)) return which ? 1.0f : y;
		return sincos_poly(x, x * x, which);
	}
	double const r = x * 0x1.45F306DC9C883p+23;      // 2/pi * 2^24
	int const n = ((int32_t)r + 0x800000) >> 24;
	x = x - (double)n * 0x1.921FB54442D18p0;
	double const s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
	int const sel = n ^ which;
	float const v = sincos_poly(x * s, x * x, sel);
	return ((n & 2) && (sel & 1)) ? -v : v;
}

// ---- Eigen::SelfAdjointEigenSolver<Matrix3f>::computeDirect --------------------------------------------------------
// Eigen's fixed-size reductions of three terms (trace, squaredNorm, dot, lazy 3x3 products): a + (b + c)
__device__ __forceinline__ float sum3(float a, float b, float c) { return a + (b + c); }
__device__ __forceinline__ float sqnorm3(f3 a) { return sum3(a.x * a.x, a.y * a.y, a.z * a.z); }
__device__ __forceinline__ f3 cross3(f3 a, f3 b)
{
	return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// symmetric 3x3, lower triangle (m10 = m(1,0) ...)
struct Sym3 { float m00, m10, m11, m20, m21, m22; };

// computeRoots (SelfAdjointEigenSolver.h:594-631)
__device__ __forceinline__ void eig3_roots(const Sym3& s, float roots[3])
{
	float const s_inv3 = 1.0f / 3.0f;
	float const s_sqrt3 = 1.7320508075688772f;   // sqrt(3.0f) rounded to float
	float const m00 = s.m00, m11 = s.m11, m22 = s.m22, m10 = s.m10, m20 = s.m20, m21 = s.m21;
	float const c0 = m00 * m11 * m22 + 2.0f * m10 * m20 * m21 - m00 * m21 * m21 - m11 * m20 * m20 - m22 * m10 * m10;
	float const c1 = m00 * m11 - m10 * m10 + m00 * m22 - m20 * m20 + m11 * m22 - m21 * m21;
	float const c2 = m00 + m11 + m22;
	float const c2_over_3 = c2 * s_inv3;
	float a_over_3 = (c2 * c2_over_3 - c1) * s_inv3;
	a_over_3 = a_over_3 < 0.0f ? 0.0f : a_over_3;
	float const half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
	float q = a_over_3 * a_over_3 * a_over_3 - half_b * half_b;
	q = q < 0.0f ? 0.0f : q;
	float const rho = sqrtf(a_over_3);
	float const theta = atan2f_fdlibm(sqrtf(q), half_b) * s_inv3;
	float const cos_theta = sincosf_glibc(theta, 1);
	float const sin_theta = sincosf_glibc(theta, 0);
	roots[0] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
	roots[1] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
	roots[2] = c2_over_3 + 2.0f * rho * cos_theta;
}

// extract_kernel (SelfAdjointEigenSolver.h:633-654); the matrix is symmetric, so column i == row i
__device__ __forceinline__ void eig3_extract_kernel(const Sym3& m, f3& res, f3& representative)
{
	f3 const col0 = mk3(m.m00, m.m10, m.m20), col1 = mk3(m.m10, m.m11, m.m21), col2 = mk3(m.m20, m.m21, m.m22);
	int i0 = 0;
	float best = fabsf(m.m00);
	if (fabsf(m.m11) > best) { best = fabsf(m.m11); i0 = 1; }
	if (fabsf(m.m22) > best) { i0 = 2; }
	f3 const a = i0 == 0 ? col0 : (i0 == 1 ? col1 : col2);      // col(i0)
	f3 const b = i0 == 0 ? col1 : (i0 == 1 ? col2 : col0);      // col((i0 + 1) % 3)
	f3 const c = i0 == 0 ? col2 : (i0 == 1 ? col0 : col1);      // col((i0 + 2) % 3)
	representative = a;
	f3 const c0 = cross3(a, b);
	float const n0 = sqnorm3(c0);
	f3 const c1 = cross3(a, c);
	float const n1 = sqnorm3(c1);
	if (n0 > n1)
	{
		float const s = sqrtf(n0);
		res = mk3(c0.x / s, c0.y / s, c0.z / s);
	}
	else
	{
		float const s = sqrtf(n1);
		res = mk3(c1.x / s, c1.y / s, c1.z / s);
	}
}

// direct_selfadjoint_eigenvalues<Solver,3,false>::run (SelfAdjointEigenSolver.h:656-740).  A: lower triangle of the
// input.  evals ascending; v0, v1, v2 = the eigenvector columns.
__device__ __forceinline__ void eig3(const Sym3& A, float evals[3], f3& v0, f3& v1, f3& v2)
{
	float const eps = 1.1920928955078125e-07f;
	float const shift = sum3(A.m00, A.m11, A.m22) / 3.0f;
	Sym3 S = A;
	S.m00 -= shift; S.m11 -= shift; S.m22 -= shift;
	float scale = fabsf(S.m00);
	if (fabsf(S.m10) > scale) scale = fabsf(S.m10);
	if (fabsf(S.m20) > scale) scale = fabsf(S.m20);
	if (fabsf(S.m11) > scale) scale = fabsf(S.m11);
	if (fabsf(S.m21) > scale) scale = fabsf(S.m21);
	if (fabsf(S.m22) > scale) scale = fabsf(S.m22);
	if (scale > 0.0f)
	{
		S.m00 /= scale; S.m10 /= scale; S.m11 /= scale; S.m20 /= scale; S.m21 /= scale; S.m22 /= scale;
	}
	eig3_roots(S, evals);
	if ((evals[2] - evals[0]) <= eps)
	{
		v0 = mk3(1.0f, 0.0f, 0.0f); v1 = mk3(0.0f, 1.0f, 0.0f); v2 = mk3(0.0f, 0.0f, 1.0f);
	}
	else
	{
		float d0 = evals[2] - evals[1];
		float const d1 = evals[1] - evals[0];
		bool const k_is_2 = d0 > d1;                 // k = 2, l = 0 when true; k = 0, l = 2 otherwise
		if (k_is_2) d0 = d1;
		float const ek = k_is_2 ? evals[2] : evals[0];
		float const el = k_is_2 ? evals[0] : evals[2];
		Sym3 tmp = S;
		tmp.m00 -= ek; tmp.m11 -= ek; tmp.m22 -= ek;
		f3 vk, vl;
		eig3_extract_kernel(tmp, vk, vl);
		if (d0 <= 2.0f * eps * d1)
		{
			float const d = sum3(vk.x * vl.x, vk.y * vl.y, vk.z * vl.z);
			vl = mk3(vl.x - d * vl.x, vl.y - d * vl.y, vl.z - d * vl.z);
			float const z = sqnorm3(vl);
			if (z > 0.0f)
			{
				float const s = sqrtf(z);
				vl = mk3(vl.x / s, vl.y / s, vl.z / s);
			}
		}
		else
		{
			tmp = S;
			tmp.m00 -= el; tmp.m11 -= el; tmp.m22 -= el;
			f3 dummy;
			eig3_extract_kernel(tmp, vl, dummy);
		}
		v0 = k_is_2 ? vl : vk;
		v2 = k_is_2 ? vk : vl;
		f3 n = cross3(v2, v0);
		float const z = sqnorm3(n);
		if (z > 0.0f)
		{
			float const s = sqrtf(z);
			n = mk3(n.x / s, n.y / s, n.z / s);
		}
		v1 = n;
	}
	evals[0] *= scale; evals[1] *= scale; evals[2] *= scale;
	evals[0] += shift; evals[1] += shift; evals[2] += shift;
}

// ---- WPCA tail: eigenvalue clamps (eq. 15 [YT13], RayMarcher.cpp:227-238) and G (:242-250) ---------------------------
struct Settings { float k_n, k_r, k_s; uint32_t n_eps; };

// glm::mat3 G, column-major: g[c * 3 + r]
struct Mat3 { float g[9]; };

__device__ __forceinline__ void wpca_G(const Sym3& C, uint32_t N, const Settings& s, float particle_radius_inv, Mat3& G)
{
	float Sigma[3];
	f3 r0, r1, r2;
	eig3(C, Sigma, r0, r1, r2);
	if (N <= s.n_eps)
	{
		Sigma[0] = Sigma[1] = Sigma[2] = s.k_n;
	}
	else
	{
		float const m12 = Sigma[1] < Sigma[2] ? Sigma[2] : Sigma[1];     // std::max
		float const mx = Sigma[0] < m12 ? m12 : Sigma[0];
		float const minimum = mx / s.k_r;
#pragma unroll
		for (int k = 0; k < 3; k++) Sigma[k] = Sigma[k] < minimum ? minimum : Sigma[k];
#pragma unroll
		for (int k = 0; k < 3; k++) Sigma[k] *= s.k_s;
	}
	// G = ParticleRadiusInv * R * Sigma^-1 * R^T: T(i,k) = (hinv * R(i,k)) * (1 / Sigma(k)); G(i,j) = sum_k T(i,k) * R(j,k)
	float const R[3][3] = { { r0.x, r1.x, r2.x }, { r0.y, r1.y, r2.y }, { r0.z, r1.z, r2.z } };   // R[i][k]
	float inv[3], T[3][3];
#pragma unroll
	for (int k = 0; k < 3; k++) inv[k] = 1.0f / Sigma[k];
#pragma unroll
	for (int i = 0; i < 3; i++)
#pragma unroll
		for (int k = 0; k < 3; k++) T[i][k] = (particle_radius_inv * R[i][k]) * inv[k];
#pragma unroll
	for (int i = 0; i < 3; i++)
#pragma unroll
		for (int j = 0; j < 3; j++) G.g[j * 3 + i] = sum3(T[i][0] * R[j][0], T[i][1] * R[j][1], T[i][2] * R[j][2]);
}

// glm::determinant(mat3) (vendor/glm/glm/detail/func_matrix.inl:211-220)
__device__ __forceinline__ float det3(const Mat3& G)
{
	const float* m = G.g;      // m[c * 3 + r]
	return +m[0] * (m[4] * m[8] - m[7] * m[5])
		- m[3] * (m[1] * m[8] - m[7] * m[2])
		+ m[6] * (m[1] * m[5] - m[4] * m[2]);
}

// glm mat3 * vec3 (vendor/glm/glm/detail/type_mat3x3.inl:468-474)
__device__ __forceinline__ f3 mat3_mul(const Mat3& G, f3 v)
{
	const float* m = G.g;
	return mk3(m[0] * v.x + m[3] * v.y + m[6] * v.z,
			   m[1] * v.x + m[4] * v.y + m[7] * v.z,
			   m[2] * v.x + m[5] * v.y + m[8] * v.z);
}

// AnisotropicKernel (Kernel.cpp:55-61): h, h^2, 1/h, sig = 8/pi
struct Kernel { float h, h_squared, h_inv, sig; };

// AnisotropicKernel::W (Kernel.cpp:63-82)
__device__ __forceinline__ float W(const Kernel& k, const Mat3& G, float detG, f3 r_)
{
	f3 const r = mat3_mul(G, r_);
	float q = (r.x * r.x + r.y * r.y) + r.z * r.z;
	if (q >= k.h_squared) return 0.0f;
	q = sqrtf(q) * k.h_inv;
	if (q >= 0.5f)
	{
		float const q_ = 1.0f - q;
		return k.sig * detG * (2.0f * q_ * q_ * q_);
	}
	return k.sig * detG * (6.0f * (q * q * q - q * q) + 1.0f);
}

// AnisotropicKernel::gradW (Kernel.cpp:84-107)
__device__ __forceinline__ f3 gradW(const Kernel& k, const Mat3& G, float detG, f3 r_)
{
	f3 const r = mat3_mul(G, r_);
	float const rn = (r.x * r.x + r.y * r.y) + r.z * r.z;
	if (rn >= k.h_squared) return mk3(0.0f, 0.0f, 0.0f);
	float const r_length = sqrtf(rn);
	float const q = r_length * k.h_inv;
	float const inv_len = 1.0f / sqrtf(rn);                       // glm::normalize(r) = r * inversesqrt(dot(r, r))
	f3 const nr = mk3(r.x * inv_len, r.y * inv_len, r.z * inv_len);
	float const den = r_length * k.h;
	f3 const gradQ = mk3(nr.x / den, nr.y / den, nr.z / den);
	if (q >= 0.5f)
	{
		float const q_ = 1.0f - q;
		float const a = -k.sig * detG;
		float const s = 6.0f * q_ * q_;
		return mk3((a * gradQ.x) * s, (a * gradQ.y) * s, (a * gradQ.z) * s);
	}
	float const a = k.sig * detG;
	float const s = 6.0f * (3.0f * q * q - 2.0f * q);
	return mk3((a * gradQ.x) * s, (a * gradQ.y) * s, (a * gradQ.z) * s);
}

// CubicKernel::W (Kernel.cpp:117-125) given d2 = dot(r, r)
__device__ __forceinline__ float cubic_W(float h, float h_inv, float d2)
{
	float const d = sqrtf(d2);
	if (d >= h) return 0.0f;
	float const k = d * h_inv;
	return 1.0f - k * k * k;
}

}  // namespace aniso
}  // namespace fm
