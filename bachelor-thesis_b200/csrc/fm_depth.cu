// fm_depth.cu -- depth pre-pass that seeds each ray's start distance (sm_100a).
//
// Replaces the reference's rasterised pass: AdvancedRenderer::CollectRenderData
// (src/app/AdvancedRenderer/AdvancedRenderer.cpp:447-485: one camera-facing quad p +- h*System[0]
// +- h*System[1] per particle, UV in [-1,1]^2), DepthRenderPass (DepthRenderPass.cpp:45-87: clear 1.0,
// cull none, depth test Less) and assets/shaders/advanced/depth.{vert,frag}
// (fragment: l2 = dot(uv,uv); discard if l2 > 1; off = cos(pi/2 sqrt(l2));
//  depth = (P * (viewPos - h*(0,0,off))).z / w).
//
// Per-fragment arithmetic (identical, op for op, to fo_depth_prepass in oracle/fluid_oracle.c): the quad lies
// in the plane view-z = z_c, so at the pixel centre (px+.5, py+.5) the view ray meets it at
// x_v = ndc_x*z_c/P00, y_v = ndc_y*z_c/P11 and uv = ((x_v - x_c)/h, (y_v - y_c)/h), folded into
// u = ndc_x*ax - bx with ax = z_c/(P00*h), bx = x_c/h.  The image is the minimum over all fragments, which
// is order-independent, so ANY superset of the winning fragments gives the same bits.  The kernels below
// only decide which (particle, pixel) pairs are worth evaluating:
//
//   k_depth_clear    depth <- 1.0 (DepthRenderPass.cpp:54); tile bounds <- 1.0
//   k_depth_bounds   hierarchical-Z seed.  Screen tiles of T x T pixels.  A particle whose disc contains all
//                    pixel centres of a tile bounds the final depth of every pixel of that tile from above by
//                    its own fragment depth at the tile's farthest corner (l2 is convex and, in FP32, monotone
//                    along each axis, so the corner maximum bounds every pixel of the tile exactly; a few ulps of
//                    slack cover the polynomial cosine).  atomicMin into tile_bound[].
//   k_depth_splat    one thread per particle tests the particle's nearest possible depth (disc centre,
//                    z_c - h) against the bounds of the tiles its pixel box overlaps; a particle that cannot win
//                    any tile (every interior particle of the fluid) costs a handful of L1/L2-resident loads.
//                    Survivors are handed to the whole warp: lanes re-test the tiles in parallel and then
//                    evaluate the fragments of the surviving tiles, 32 pixels at a time, atomicMin on the
//                    raw bits (depth in [0,1] orders like its uint bits).
//
// Cost at C2 (1M particles, 1080p): ~1.05 G warp instructions for the brute-force splat -> see profiles/.
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
	int tiles_x, tiles_y;
};

// everything the fragment evaluation needs about one particle
struct Splat
{
	float z_c, ax, bx, ay, by;
	uint32_t near_bits;          // depth bits of the nearest fragment this particle can produce
	int x0, y0, x1, y1;          // conservative pixel box, clipped to the screen (x1 < x0: nothing to draw)
};

__global__ void __launch_bounds__(256) k_depth_clear(uint32_t* __restrict__ depth_bits, uint32_t npix,
													 uint32_t* __restrict__ tile_bound, uint32_t ntiles)
{
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < npix) depth_bits[i] = 0x3f800000u;   // 1.0f: DepthRenderPass.cpp:54
	if (i < ntiles) tile_bound[i] = 0x3f800000u;
}

__device__ __forceinline__ float frag_depth(const DepthParams& dp, float z_c, float l2)
{
	float const off = cos_half_pi(sqrtr(l2));           // depth.frag:25
	float const zf = subr(z_c, mulr(dp.h, off));        // depth.frag:28
	float d = divr(addr(mulr(dp.P22, zf), dp.P32), zf); // depth.frag:30-33
	return d < 0.0f ? 0.0f : (d > 1.0f ? 1.0f : d);
}

__device__ __forceinline__ float frag_u(float pix, float two_inv, float a, float b)
{
	float const ndc = subr(mulr(addr(pix, 0.5f), two_inv), 1.0f);
	return subr(mulr(ndc, a), b);
}

__device__ __forceinline__ bool splat_setup(const DepthParams& dp, float4 p, Splat& s)
{
	s.x1 = -1; s.x0 = 0; s.y0 = 0; s.y1 = -1;
	// depth.vert:20-27: viewPosition = View * vec4(p, 1); ViewPosition = xyz / w
	float vc[4];
	mat4_mul_vec4(dp.view, p.x, p.y, p.z, 1.0f, vc);
	float const x_c = divr(vc[0], vc[3]), y_c = divr(vc[1], vc[3]), z_c = divr(vc[2], vc[3]);
	if (!(z_c > 0.0f)) return false;
	float const quad_depth = divr(addr(mulr(dp.P22, z_c), dp.P32), z_c);
	if (!(quad_depth >= 0.0f && quad_depth <= 1.0f)) return false;   // quad clipped by near / far

	// conservative pixel box of the disc (any superset yields the same image)
	float const cx = mulr(addr(divr(mulr(dp.P00, x_c), z_c), 1.0f), dp.half_w);
	float const cy = mulr(addr(divr(mulr(dp.P11, y_c), z_c), 1.0f), dp.half_h);
	float const rx = addr(mulr(divr(mulr(fabsf(dp.P00), dp.h), z_c), dp.half_w), 1.0f);
	float const ry = addr(mulr(divr(mulr(fabsf(dp.P11), dp.h), z_c), dp.half_h), 1.0f);
	float const fx0 = floorf(cx - rx - 0.5f), fx1 = ceilf(cx + rx - 0.5f);
	float const fy0 = floorf(cy - ry - 0.5f), fy1 = ceilf(cy + ry - 0.5f);
	if (!(fx1 >= 0.0f && fy1 >= 0.0f && fx0 <= (float)(dp.W - 1) && fy0 <= (float)(dp.H - 1))) return false;
	s.x0 = fx0 < 0.0f ? 0 : (int)fx0;
	s.y0 = fy0 < 0.0f ? 0 : (int)fy0;
	s.x1 = fx1 > (float)(dp.W - 1) ? dp.W - 1 : (int)fx1;
	s.y1 = fy1 > (float)(dp.H - 1) ? dp.H - 1 : (int)fy1;

	s.z_c = z_c;
	s.ax = divr(z_c, mulr(dp.P00, dp.h)); s.bx = mulr(x_c, dp.h_inv);
	s.ay = divr(z_c, mulr(dp.P11, dp.h)); s.by = mulr(y_c, dp.h_inv);
	// nearest depth this particle can produce: fragment at the disc centre, zf = z_c - h
	float const zn = subr(z_c, dp.h);
	float dnear = divr(addr(mulr(dp.P22, zn), dp.P32), zn);
	dnear = (zn > 0.0f && dnear > 0.0f) ? dnear : 0.0f;
	s.near_bits = __float_as_uint(dnear);
	return true;
}

template <int T>
__global__ void __launch_bounds__(256) k_depth_bounds(const float4* __restrict__ sorted, uint32_t n, DepthParams dp,
													  uint32_t* __restrict__ tile_bound)
{
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	Splat s;
	if (!splat_setup(dp, __ldg(sorted + i), s)) return;
	int const tx0 = s.x0 / T, tx1 = s.x1 / T, ty0 = s.y0 / T, ty1 = s.y1 / T;
	for (int ty = ty0; ty <= ty1; ty++)
	{
		int const py0 = ty * T, py1 = min(py0 + T - 1, dp.H - 1);
		float const v0 = frag_u((float)py0, dp.two_h_inv, s.ay, s.by);
		float const v1 = frag_u((float)py1, dp.two_h_inv, s.ay, s.by);
		float const vv = fmaxf(mulr(v0, v0), mulr(v1, v1));
		if (vv > 1.0f) continue;
		for (int tx = tx0; tx <= tx1; tx++)
		{
			int const px0 = tx * T, px1 = min(px0 + T - 1, dp.W - 1);
			float const u0 = frag_u((float)px0, dp.two_w_inv, s.ax, s.bx);
			float const u1 = frag_u((float)px1, dp.two_w_inv, s.ax, s.bx);
			// largest l2 any pixel of the tile evaluates to (FP32 mul/add are monotone)
			float const l2 = addr(fmaxf(mulr(u0, u0), mulr(u1, u1)), vv);
			if (l2 > 1.0f) continue;                          // some pixel of the tile is discarded: no bound
			float const d = frag_depth(dp, s.z_c, l2);
			uint32_t const bound = __float_as_uint(d) + 8u;   // slack for the non-monotone last bits of the cosine
			if (bound >= 0x3f800000u) continue;
			uint32_t* const cell = tile_bound + (size_t)ty * dp.tiles_x + tx;
			if (bound < __ldcg(cell)) atomicMin(cell, bound);
		}
	}
}

template <int T>
__global__ void __launch_bounds__(256) k_depth_splat(const float4* __restrict__ sorted, uint32_t n, DepthParams dp,
													 const uint32_t* __restrict__ tile_bound,
													 uint32_t* __restrict__ depth_bits)
{
	constexpr int PIX = T * T;                       // pixels per tile
	constexpr int LPT = PIX < 32 ? PIX : 32;         // lanes working on one tile
	constexpr int TPR = 32 / LPT;                    // tiles per round
	constexpr int RPT = PIX / LPT;                   // rounds per tile
	uint32_t const lane = threadIdx.x & 31u;
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;

	Splat s;
	bool live = false;
	if (i < n) live = splat_setup(dp, __ldg(sorted + (dp.reverse ? (n - 1u - i) : i)), s);
	int tx0 = 0, ty0 = 0, ntx = 0, nty = 0;
	bool wins = false;
	if (live)
	{
		tx0 = s.x0 / T; ty0 = s.y0 / T;
		ntx = s.x1 / T - tx0 + 1; nty = s.y1 / T - ty0 + 1;
		// can this particle still win a pixel of any tile it overlaps?
		for (int ty = 0; ty < nty && !wins; ty++)
		{
			const uint32_t* row = tile_bound + (size_t)(ty0 + ty) * dp.tiles_x + tx0;
			for (int tx = 0; tx < ntx; tx++)
				if (s.near_bits < __ldg(row + tx)) { wins = true; break; }
		}
	}

	// survivors are processed by the whole warp, one particle at a time
	uint32_t todo = __ballot_sync(0xffffffffu, wins);
	while (todo)
	{
		int const src = __ffs(todo) - 1;
		todo &= todo - 1u;
		float const z_c = __shfl_sync(0xffffffffu, s.z_c, src);
		float const ax = __shfl_sync(0xffffffffu, s.ax, src), bx = __shfl_sync(0xffffffffu, s.bx, src);
		float const ay = __shfl_sync(0xffffffffu, s.ay, src), by = __shfl_sync(0xffffffffu, s.by, src);
		uint32_t const near_bits = __shfl_sync(0xffffffffu, s.near_bits, src);
		int const bx0 = __shfl_sync(0xffffffffu, s.x0, src), bx1 = __shfl_sync(0xffffffffu, s.x1, src);
		int const by0 = __shfl_sync(0xffffffffu, s.y0, src), by1 = __shfl_sync(0xffffffffu, s.y1, src);
		int const ptx0 = bx0 / T, pty0 = by0 / T;
		int const pntx = bx1 / T - ptx0 + 1;
		int const ntiles = pntx * (by1 / T - pty0 + 1);
		float const inv_ntx = 1.0f / (float)pntx;

		for (int tb = 0; tb < ntiles; tb += 32)
		{
			int const t = tb + (int)lane;
			// t / pntx for the small ints involved
			int trow = (int)(((float)t + 0.5f) * inv_ntx);
			int tcol = t - trow * pntx;
			if (tcol < 0) { trow--; tcol += pntx; } else if (tcol >= pntx) { trow++; tcol -= pntx; }
			bool const alive = t < ntiles && near_bits < __ldg(tile_bound + (size_t)(pty0 + trow) * dp.tiles_x + (ptx0 + tcol));
			uint32_t tiles = __ballot_sync(0xffffffffu, alive);
			int const my_tile_xy = ((pty0 + trow) << 16) | (ptx0 + tcol);

			while (tiles)
			{
				// lane group g takes the g-th surviving tile of this batch
				uint32_t const sel = __fns(tiles, 0, (int)(lane / LPT) + 1);
#pragma unroll
				for (int k = 0; k < TPR; k++) tiles &= tiles - 1u;
				int const txy = __shfl_sync(0xffffffffu, my_tile_xy, sel & 31u);
				if (sel == 0xffffffffu) continue;
				int const tpx = (txy & 0xffff) * T, tpy = (txy >> 16) * T;
#pragma unroll
				for (int r = 0; r < RPT; r++)
				{
					int const k = r * LPT + (int)(lane % LPT);
					int const px = tpx + (k % T), py = tpy + (k / T);
					if (px < bx0 || px > bx1 || py < by0 || py > by1) continue;
					uint32_t* const cell = depth_bits + (size_t)py * (size_t)dp.W + (size_t)px;
					if (near_bits >= __ldcg(cell)) continue;              // cannot win this pixel (L2 read: always fresh)
					float const u = frag_u((float)px, dp.two_w_inv, ax, bx);
					float const v = frag_u((float)py, dp.two_h_inv, ay, by);
					float const l2 = addr(mulr(u, u), mulr(v, v));
					if (l2 > 1.0f) continue;                            // depth.frag:22 `if (l2 > 1) discard;`
					float const d = frag_depth(dp, z_c, l2);
					if (!(d < 1.0f)) continue;                          // compare Less against the clear value
					atomicMin(cell, __float_as_uint(d));
				}
			}
		}
	}
}

template <int T>
void launch_tiles(Context* ctx, const Frame& f, DepthParams dp, uint32_t* tile_bound)
{
	uint32_t const n = (uint32_t)f.n;
	uint32_t const blocks = (n + 255u) / 256u;
	k_depth_bounds<T><<<blocks, 256, 0, ctx->stream>>>(f.d_sorted, n, dp, tile_bound);
	k_depth_splat<T><<<blocks, 256, 0, ctx->stream>>>(f.d_sorted, n, dp, tile_bound, (uint32_t*)ctx->d_depth);
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
	float const adx = fabsf(d[0]), ady = fabsf(d[1]), adz = fabsf(d[2]);
	float const dom = (adz >= adx && adz >= ady) ? d[2] : (adx >= ady ? d[0] : d[1]);
	dp.reverse = dom < 0.0f ? 1 : 0;

	// tile size from the projected disc radius at the centre of the particle AABB (a tuning choice only:
	// every T gives the same image)
	float c[3], vz = 0.0f, vw = 0.0f;
	for (int a = 0; a < 3; a++) c[a] = 0.5f * (f.gp.mn[a] + f.gp.mx[a]);
	vz = cam.view[2] * c[0] + cam.view[6] * c[1] + cam.view[10] * c[2] + cam.view[14];
	vw = cam.view[3] * c[0] + cam.view[7] * c[1] + cam.view[11] * c[2] + cam.view[15];
	float const zc = (vw != 0.0f) ? vz / vw : vz;
	float const r_px = (zc > f.h) ? fabsf(dp.P11) * f.h / zc * dp.half_h : 1e9f;
	int const T = r_px < 5.0f ? 2 : (r_px < 12.0f ? 4 : (r_px < 28.0f ? 8 : 16));
	dp.tiles_x = (ctx->width + T - 1) / T;
	dp.tiles_y = (ctx->height + T - 1) / T;
	uint32_t const ntiles = (uint32_t)dp.tiles_x * (uint32_t)dp.tiles_y;
	int rc;
	if ((rc = ensure_capacity(&ctx->d_tile_bound, &ctx->cap_tile_bound, (size_t)ntiles))) return rc;

	uint32_t const npix = (uint32_t)ctx->width * (uint32_t)ctx->height;
	k_depth_clear<<<(npix + 255) / 256, 256, 0, ctx->stream>>>((uint32_t*)ctx->d_depth, npix, ctx->d_tile_bound, ntiles);
	switch (T)
	{
	case 2: launch_tiles<2>(ctx, f, dp, ctx->d_tile_bound); break;
	case 4: launch_tiles<4>(ctx, f, dp, ctx->d_tile_bound); break;
	case 8: launch_tiles<8>(ctx, f, dp, ctx->d_tile_bound); break;
	default: launch_tiles<16>(ctx, f, dp, ctx->d_tile_bound); break;
	}
	ctx->kernel_launches += 3;
	FM_CUDA(cudaGetLastError());
	return FR_OK;
}

}  // namespace fm
