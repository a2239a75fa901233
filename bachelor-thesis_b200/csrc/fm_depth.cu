// fm_depth.cu -- depth pre-pass that seeds each ray's start distance (sm_100a).
//
// Replaces the reference's rasterised pass: AdvancedRenderer::CollectRenderData
// (src/app/AdvancedRenderer/AdvancedRenderer.cpp:447-485: one camera-facing quad p +- h*System[0]
// +- h*System[1] per particle, UV in [-1,1]^2), DepthRenderPass (DepthRenderPass.cpp:45-87: clear 1.0,
// cull none, depth test Less) and assets/shaders/advanced/depth.{vert,frag}
// (fragment: l2 = dot(uv,uv); discard if l2 > 1; off = cos(pi/2 sqrt(l2));
//  depth = (P * (viewPos - h*(0,0,off))).z / w).
//
// Analytic form (identical, op for op, to fo_depth_prepass in oracle/fluid_oracle.c): the quad lies in
// the plane view-z = z_c, so at the pixel centre (px+.5, py+.5) the view ray meets it at
// x_v = ndc_x*z_c/P00, y_v = ndc_y*z_c/P11 and uv = ((x_v - x_c)/h, (y_v - y_c)/h), folded into
// u = ndc_x*ax - bx with ax = z_c/(P00*h), bx = x_c/h (one multiply-subtract per pixel).
// The minimum over particles is order-independent, so the image is bit-reproducible.
//
// Mapping: one warp per particle, lanes sweep the particle's pixel box; a fragment is evaluated
// fully only if the particle's nearest possible depth (at z_c - h) beats the stored depth, so occluded
// particles cost one L2-resident load + compare per pixel.  Depth in [0,1) orders like its uint bits
// -> atomicMin on the raw bits.
#include "fm_internal.h"

namespace fm
{

namespace
{

struct DepthParams
{
	float view[16];
	float P00, P11, P22, P32;
	float h, h_inv;
	int W, H;
	float two_w_inv, two_h_inv, half_w, half_h;
	int reverse;
};

__global__ void __launch_bounds__(256) k_depth_clear(uint32_t* __restrict__ depth_bits, uint32_t npix)
{
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < npix) depth_bits[i] = 0x3f800000u;   // 1.0f: DepthRenderPass.cpp:54
}

__global__ void __launch_bounds__(256) k_depth_splat(const float4* __restrict__ sorted, uint32_t n, DepthParams dp,
													 uint32_t* __restrict__ depth_bits)
{
	uint32_t const lane = threadIdx.x & 31u;
	uint32_t const warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += warps)
	{
		uint32_t const i = dp.reverse ? (n - 1u - w) : w;
		float4 const p = __ldg(sorted + i);
		// depth.vert:20-27: viewPosition = View * vec4(p, 1); ViewPosition = xyz / w
		float vc[4];
		mat4_mul_vec4(dp.view, p.x, p.y, p.z, 1.0f, vc);
		float const x_c = divr(vc[0], vc[3]), y_c = divr(vc[1], vc[3]), z_c = divr(vc[2], vc[3]);
		if (!(z_c > 0.0f)) continue;
		float const quad_depth = divr(addr(mulr(dp.P22, z_c), dp.P32), z_c);
		if (!(quad_depth >= 0.0f && quad_depth <= 1.0f)) continue;   // quad clipped by near / far

		// conservative pixel box of the disc (any superset yields the same image)
		float const cx = mulr(addr(divr(mulr(dp.P00, x_c), z_c), 1.0f), dp.half_w);
		float const cy = mulr(addr(divr(mulr(dp.P11, y_c), z_c), 1.0f), dp.half_h);
		float const rx = addr(mulr(divr(mulr(fabsf(dp.P00), dp.h), z_c), dp.half_w), 1.0f);
		float const ry = addr(mulr(divr(mulr(fabsf(dp.P11), dp.h), z_c), dp.half_h), 1.0f);
		float const fx0 = floorf(cx - rx - 0.5f), fx1 = ceilf(cx + rx - 0.5f);
		float const fy0 = floorf(cy - ry - 0.5f), fy1 = ceilf(cy + ry - 0.5f);
		if (!(fx1 >= 0.0f && fy1 >= 0.0f && fx0 <= (float)(dp.W - 1) && fy0 <= (float)(dp.H - 1))) continue;
		int const x0 = fx0 < 0.0f ? 0 : (int)fx0;
		int const y0 = fy0 < 0.0f ? 0 : (int)fy0;
		int const x1 = fx1 > (float)(dp.W - 1) ? dp.W - 1 : (int)fx1;
		int const y1 = fy1 > (float)(dp.H - 1) ? dp.H - 1 : (int)fy1;
		int const bw = x1 - x0 + 1;
		int const total = bw * (y1 - y0 + 1);

		float const ax = divr(z_c, mulr(dp.P00, dp.h)), bx = mulr(x_c, dp.h_inv);
		float const ay = divr(z_c, mulr(dp.P11, dp.h)), by = mulr(y_c, dp.h_inv);
		// nearest depth this particle can produce: fragment at the disc centre, zf = z_c - h
		float const zn = subr(z_c, dp.h);
		float dnear = divr(addr(mulr(dp.P22, zn), dp.P32), zn);
		dnear = (zn > 0.0f && dnear > 0.0f) ? dnear : 0.0f;
		uint32_t const near_bits = __float_as_uint(dnear);
		float const inv_bw = 1.0f / (float)bw;

		for (int j = (int)lane; j < total; j += 32)
		{
			int row = (int)(((float)j + 0.5f) * inv_bw);       // j / bw for the small ints involved
			int col = j - row * bw;
			if (col < 0) { row--; col += bw; } else if (col >= bw) { row++; col -= bw; }
			int const px = x0 + col, py = y0 + row;
			uint32_t* const cell = depth_bits + (size_t)py * (size_t)dp.W + (size_t)px;
			if (near_bits >= __ldcg(cell)) continue;              // cannot win this pixel (L2 read: always fresh)
			float const ndc_x = subr(mulr(addr((float)px, 0.5f), dp.two_w_inv), 1.0f);
			float const ndc_y = subr(mulr(addr((float)py, 0.5f), dp.two_h_inv), 1.0f);
			float const u = subr(mulr(ndc_x, ax), bx);
			float const v = subr(mulr(ndc_y, ay), by);
			float const l2 = addr(mulr(u, u), mulr(v, v));
			if (l2 > 1.0f) continue;                            // depth.frag:22 `if (l2 > 1) discard;`
			float const off = cos_half_pi(sqrtr(l2));           // depth.frag:25
			float const zf = subr(z_c, mulr(dp.h, off));        // depth.frag:28
			float d = divr(addr(mulr(dp.P22, zf), dp.P32), zf); // depth.frag:30-33
			d = d < 0.0f ? 0.0f : (d > 1.0f ? 1.0f : d);
			if (!(d < 1.0f)) continue;                          // compare Less against the clear value
			atomicMin(cell, __float_as_uint(d));
		}
	}
}

}  // namespace

int launch_depth_prepass(Context* ctx, const Frame& f)
{
	const fr_camera& cam = ctx->camera;
	// only the perspective structure glm::perspectiveLH_ZO scaled by (1,-1,1) produces is supported
	// (src/engine/camera/Camera3D.cpp:11, vendor/glm/glm/ext/matrix_clip_space.inl:265-278)
	static const int zero_idx[] = { 1, 2, 3, 4, 6, 7, 8, 9, 12, 13, 15 };
	for (int k : zero_idx)
		if (cam.projection[k] != 0.0f) { set_error("depth pre-pass: projection is not a glm perspective matrix"); return FR_ERR_INVALID; }
	if (cam.projection[11] != 1.0f) { set_error("depth pre-pass: projection[2][3] must be 1 (left-handed perspective)"); return FR_ERR_INVALID; }

	DepthParams dp;
	for (int k = 0; k < 16; k++) dp.view[k] = cam.view[k];
	dp.P00 = cam.projection[0]; dp.P11 = cam.projection[5]; dp.P22 = cam.projection[10]; dp.P32 = cam.projection[14];
	dp.h = f.h;
	dp.h_inv = 1.0f / f.h;
	dp.W = ctx->width; dp.H = ctx->height;
	dp.two_w_inv = 2.0f / (float)ctx->width;
	dp.two_h_inv = 2.0f / (float)ctx->height;
	dp.half_w = 0.5f * (float)ctx->width;
	dp.half_h = 0.5f * (float)ctx->height;
	// particles are stored x-major / z-fastest: walk them so that the nearer ones tend to come first,
	// which lets the `cannot win` test reject most occluded fragments
	const float* d = cam.direction;
	float const ax = fabsf(d[0]), ay = fabsf(d[1]), az = fabsf(d[2]);
	float const dom = (az >= ax && az >= ay) ? d[2] : (ax >= ay ? d[0] : d[1]);
	dp.reverse = dom < 0.0f ? 1 : 0;

	uint32_t const npix = (uint32_t)ctx->width * (uint32_t)ctx->height;
	cudaStream_t const s = ctx->stream;
	k_depth_clear<<<(npix + 255) / 256, 256, 0, s>>>((uint32_t*)ctx->d_depth, npix);
	uint32_t const n = (uint32_t)f.n;
	uint64_t const want_blocks = ((uint64_t)n * 32 + 255) / 256;
	uint32_t const blocks = (uint32_t)(want_blocks < (uint64_t)ctx->sm_count * 64 ? want_blocks : (uint64_t)ctx->sm_count * 64);
	k_depth_splat<<<blocks, 256, 0, s>>>(f.d_sorted, n, dp, (uint32_t*)ctx->d_depth);
	ctx->kernel_launches += 2;
	FM_CUDA(cudaGetLastError());
	return FR_OK;
}

}  // namespace fm
