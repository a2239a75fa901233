// fm_smooth.cu -- screen-space smoothing of the depth image and the normals derived from it (sm_100a).
//
// Replaces GaussRenderPass (src/app/AdvancedRenderer/GaussRenderPass.cpp:15-66 kernel weights, assets/shaders/advanced/
// gauss.frag:28-47) -- which the reference runs every UI frame (AdvancedRenderer.cpp:217-230, s_EnableGaussPass = true) --
// and the Sobel normal of the unprojected smoothed depth (composition.frag:50-57,87-104 with the neighbour coordinates
// of assets/shaders/fullscreen.vert:30-42), which the reference has switched off (`#if 0`).  Arithmetic: FP32, one
// rounding per operation, in the shaders' loop and expression order (oracle: fo_gauss_depth / fo_sobel_normals).
// Spread = 1 (GaussRenderPass.h:66): every tap is an exact texel, addressing clamps to the edge
// (BilateralBuffer.cpp:248-250).  HBM-bound: (2N+1)^2 taps per pixel out of L1/L2, 4 B per pixel in and out.
#include "fm_internal.h"

#include <math.h>

#include <vector>

namespace fm
{

namespace
{

constexpr int kMaxGaussN = 31;            // MAX_KERNEL_N = 32 (gauss.frag:3)

__global__ void __launch_bounds__(256) k_gauss_depth(const float* __restrict__ depth, int W, int H, int n, const float* __restrict__ weights,
													  float* __restrict__ smoothed)
{
	int const x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
	if (x >= W || y >= H) return;
	float sum = 0.0f;
	for (int i = -n; i <= n; i++)
	{
		int const sx = min(max(x + i, 0), W - 1);
		for (int j = -n; j <= n; j++)
		{
			int const sy = min(max(y + j, 0), H - 1);
			float const k = __ldg(weights + abs(i) * (n + 1) + abs(j));
			sum = addr(sum, mulr(k, __ldg(depth + (size_t)sy * W + sx)));
		}
	}
	smoothed[(size_t)y * W + x] = sum;
}

struct InvProj { float m[16]; };

// smoothedPosition (composition.frag:50-57): the texel a linear sampler returns at a texel centre, unprojected
__device__ __forceinline__ f3 smoothed_position(const float* __restrict__ smoothed, int W, int H, const InvProj& ip, float u, float v)
{
	int const tx = min(max((int)floorf(mulr(u, (float)W)), 0), W - 1), ty = min(max((int)floorf(mulr(v, (float)H)), 0), H - 1);
	float h[4];
	mat4_mul_vec4(ip.m, subr(mulr(2.0f, u), 1.0f), subr(mulr(2.0f, v), 1.0f), __ldg(smoothed + (size_t)ty * W + tx), 1.0f, h);
	return divs3(mk3(h[0], h[1], h[2]), h[3]);
}

__global__ void __launch_bounds__(256) k_sobel_normals(const float* __restrict__ smoothed, int W, int H, InvProj ip, float4* __restrict__ nrm)
{
	int const x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
	if (x >= W || y >= H) return;
	float const tw = divr(1.0f, (float)W), th = divr(1.0f, (float)H);
	float const u = divr(addr((float)x, 0.5f), (float)W), v = divr(addr((float)y, 0.5f), (float)H);
	float const ul = subr(u, tw), ur = addr(u, tw), vt = subr(v, th), vb = addr(v, th);
	f3 const tl = smoothed_position(smoothed, W, H, ip, ul, vt), tm = smoothed_position(smoothed, W, H, ip, u, vt),
		tr = smoothed_position(smoothed, W, H, ip, ur, vt), ml = smoothed_position(smoothed, W, H, ip, ul, v),
		mr = smoothed_position(smoothed, W, H, ip, ur, v), bl = smoothed_position(smoothed, W, H, ip, ul, vb),
		bm = smoothed_position(smoothed, W, H, ip, u, vb), br = smoothed_position(smoothed, W, H, ip, ur, vb);
	f3 const dx = sub3(sub3(sub3(add3(add3(scale3(tr, 1.0f), scale3(mr, 2.0f)), scale3(br, 1.0f)), scale3(tl, 1.0f)), scale3(ml, 2.0f)), scale3(bl, 1.0f));
	f3 const dy = sub3(sub3(sub3(add3(add3(scale3(bl, 1.0f), scale3(bm, 2.0f)), scale3(br, 1.0f)), scale3(tl, 1.0f)), scale3(tm, 2.0f)), scale3(tr, 1.0f));
	f3 const c = mk3(subr(mulr(dx.y, dy.z), mulr(dy.y, dx.z)), subr(mulr(dx.z, dy.x), mulr(dy.z, dx.x)), subr(mulr(dx.x, dy.y), mulr(dy.x, dx.y)));
	f3 const n = scale3(c, divr(1.0f, sqrtr(dot3(c, c))));
	nrm[(size_t)y * W + x] = make_float4(n.x, n.y, n.z, 1.0f);
}

}  // namespace

}  // namespace fm

using namespace fm;

extern "C" {

// GaussRenderPass.cpp:15-66: ComputeGaussKernel
int fr_gauss_kernel(int gauss_n, float* out)
{
	if (gauss_n < 0 || gauss_n > kMaxGaussN || !out) { set_error("fr_gauss_kernel: need 0 <= N <= 31"); return FR_ERR_INVALID; }
	volatile float const E = (float)2.7182818284590452353602874713527;
	float sum = 0.0f;
	for (int i = 0; i <= gauss_n; i++)
		for (int j = 0; j <= gauss_n; j++)
		{
			volatile float const x = sqrtf((float)i * (float)i + (float)j * (float)j);
			volatile float const e = -0.5f * x * x;
			float const y = powf(E, e);
			out[j + i * (gauss_n + 1)] = y;
			if (i == 0 && j == 0) sum += y;
			else if (i != 0 && j != 0) sum += 4.0f * y;
			else sum += 2.0f * y;
		}
	for (int i = 0; i <= gauss_n; i++)
		for (int j = 0; j <= gauss_n; j++) out[j + i * (gauss_n + 1)] /= sum;
	return FR_OK;
}

int fr_smooth_depth(fr_context* ctx, int gauss_n, const float* inv_projection, float* smoothed_host, float* screen_normals_host)
{
	if (!ctx) { set_error("null context"); return FR_ERR_INVALID; }
	FM_CUDA(cudaSetDevice(ctx->device));
	if (gauss_n < 0 || gauss_n > kMaxGaussN) { set_error("fr_smooth_depth: need 0 <= N <= 31 (MAX_KERNEL_N, gauss.frag:3)"); return FR_ERR_INVALID; }
	if (!ctx->have_depth) { set_error("fr_smooth_depth: no depth image (render with FR_PASS_DEPTH or call fr_set_depth first)"); return FR_ERR_STATE; }
	if (screen_normals_host && !inv_projection) { set_error("fr_smooth_depth: screen normals need Camera3D::InvProjection"); return FR_ERR_INVALID; }
	int rc = fr_wait(ctx);
	if (rc) return rc;
	int const W = ctx->width, H = ctx->height;
	size_t const npix = (size_t)W * H;
	std::vector<float> weights((size_t)(gauss_n + 1) * (gauss_n + 1));
	if ((rc = fr_gauss_kernel(gauss_n, weights.data()))) return rc;
	if ((rc = ensure_capacity(&ctx->d_smooth, &ctx->cap_smooth, npix * 5 + weights.size()))) return rc;      // depth' + normals + weights
	float* const d_smoothed = ctx->d_smooth;
	float4* const d_nrm = (float4*)(ctx->d_smooth + npix);
	float* const d_w = ctx->d_smooth + npix * 5;
	cudaStream_t const s = ctx->stream;
	FM_CUDA(cudaMemcpyAsync(d_w, weights.data(), weights.size() * 4, cudaMemcpyHostToDevice, s));
	dim3 const grid((W + 31) / 32, (H + 7) / 8);
	k_gauss_depth<<<grid, 256, 0, s>>>(ctx->d_depth, W, H, gauss_n, d_w, d_smoothed);
	ctx->kernel_launches += 1;
	if (screen_normals_host)
	{
		InvProj ip;
		for (int k = 0; k < 16; k++) ip.m[k] = inv_projection[k];
		k_sobel_normals<<<grid, 256, 0, s>>>(d_smoothed, W, H, ip, d_nrm);
		ctx->kernel_launches += 1;
		FM_CUDA(cudaMemcpyAsync(screen_normals_host, d_nrm, npix * 16, cudaMemcpyDeviceToHost, s));
	}
	FM_CUDA(cudaGetLastError());
	if (smoothed_host) FM_CUDA(cudaMemcpyAsync(smoothed_host, d_smoothed, npix * 4, cudaMemcpyDeviceToHost, s));
	return stream_sync(ctx);
}

}  // extern "C"
