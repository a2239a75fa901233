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
//   k_depth_clear    tile bounds <- 1.0; survivor count and march counters <- 0 (depth <- 1.0, DepthRenderPass.cpp:54: k_depth_seed)
//   k_depth_seed     per particle: projection, splat record (32 B), and the cheapest useful bound -- the tile under
//                    the disc centre.  Screen tiles of T x T pixels.  A particle whose disc contains all pixel
//                    centres of a tile bounds the final depth of every pixel of that tile from above by its own
//                    fragment depth at the tile's farthest corner (l2 is convex and, in FP32, monotone along each
//                    axis, so the corner maximum bounds every pixel of the tile exactly; a few ulps of slack cover
//                    the polynomial cosine).  atomicMin into tile_bound[].
//   k_depth_gate     lists the particles still in front of the bound of their own centre tile (front layers)
//   k_depth_bounds   one warp per listed particle, lanes = tiles: the bound for every tile the disc covers completely
//   k_depth_coarse   maximum of the bounds over every 4 x 4 block of tiles
//   k_depth_cull     one thread per particle tests its nearest possible depth (disc centre, z_c - h) against the
//                    coarse blocks, then the tiles, its pixel box overlaps; a particle that cannot win any tile
//                    (every interior particle of the fluid) costs a handful of L1/L2-resident loads; the rest are
//                    compacted into the survivor list (one atomic per CTA)
//   k_depth_splat    one warp per survivor: lanes re-test the tiles in parallel, compact the open ones through shared
//                    memory and evaluate their fragments 32 pixels at a time, atomicMin on the raw bits (depth in
//                    [0,1] orders like its uint bits)
//
// Cost at C2 (1M particles, 1080p): ~1.05 G warp instructions for a brute-force splat of every fragment (1.29 ms, r01a);
// ~70 M for this pipeline (0.16 ms) -> see profiles/.
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
	// tile-parallel multi-GPU (fr_set_tile_partition): only pixels of this rank's screen tiles are needed.  The
	// partition tile size is a multiple of 64 and every hi-Z tile / coarse block size divides 64, so a hi-Z tile or
	// coarse block has exactly one owner
	int part_rank, part_world, part_sw, part_sh, part_tiles_x;     // tile width / height = 1 << part_sw / part_sh
	// region partition (fr_set_region_partition): only the pixel rectangle [rx0, rx1) x [ry0, ry1) is needed; its bounds are
	// multiples of 64 (or the image edge), so a hi-Z tile or coarse block lies inside or outside as a whole.  rx1 == 0: none
	int rx0, ry0, rx1, ry1;
	int affine_view;             // the view matrix' last row is 0 0 0 1: w = 1 exactly
};

__device__ __forceinline__ bool tile_owned(const DepthParams& dp, int tx, int ty)
{
	return (uint32_t)(ty * dp.part_tiles_x + tx) % (uint32_t)dp.part_world == (uint32_t)dp.part_rank;
}

// does this rank own the partition tile / region that holds pixel (px, py)?
__device__ __forceinline__ bool pixel_owned(const DepthParams& dp, int px, int py)
{
	if (dp.rx1 > 0) return px >= dp.rx0 && px < dp.rx1 && py >= dp.ry0 && py < dp.ry1;
	if (dp.part_world <= 1) return true;
	return tile_owned(dp, px >> dp.part_sw, py >> dp.part_sh);
}

// does the pixel box touch any partition tile of this rank / its region?
__device__ __forceinline__ bool box_owned(const DepthParams& dp, int x0, int y0, int x1, int y1)
{
	if (dp.rx1 > 0) return x1 >= dp.rx0 && x0 < dp.rx1 && y1 >= dp.ry0 && y0 < dp.ry1;
	if (dp.part_world <= 1) return true;
	for (int ty = y0 >> dp.part_sh; ty <= (y1 >> dp.part_sh); ty++)
		for (int tx = x0 >> dp.part_sw; tx <= (x1 >> dp.part_sw); tx++)
			if (tile_owned(dp, tx, ty)) return true;
	return false;
}

// everything the fragment evaluation needs about one particle
struct Splat
{
	float z_c, ax, bx, ay, by;
	uint32_t near_bits;          // depth bits of the nearest fragment this particle can produce
	int x0, y0, x1, y1;          // conservative pixel box, clipped to the screen (x1 < x0: nothing to draw)
};

// (npix pixels of a rectangle rw wide at (rx0, ry0) of an image W wide: the whole image, or the context's region)
__global__ void __launch_bounds__(256) k_depth_clear(uint32_t* __restrict__ depth_bits, uint32_t npix, uint32_t rw, uint32_t rx0, uint32_t ry0, uint32_t W,
													 uint32_t* __restrict__ tile_bound, uint32_t ntiles,
													 uint32_t* __restrict__ n_survivors, uint32_t* __restrict__ counters, uint32_t counter_words)
{
	pdl_enter();
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < 4u) n_survivors[i] = 0u;             // survivor count (+ padding) of k_depth_cull
	if (i < counter_words) counters[i] = 0u;     // DeviceCounters + work-list control words of the march that follows
	if (i < npix) depth_bits[rw == W ? (size_t)i : (size_t)(ry0 + i / rw) * W + rx0 + i % rw] = 0x3f800000u;   // 1.0f: DepthRenderPass.cpp:54
	if (i < ntiles) tile_bound[i] = 0x3f800000u;
}

// particles in the sorted array: the frame's, or fewer under a region partition (k_scan_flags); none of an unusable frame
__device__ __forceinline__ uint32_t sorted_count(const GridParams* __restrict__ gp, uint32_t n)
{
	if (!gp) return n;                  // particles straight from the input array (pre-pass beside the build)
	return gp->status ? 0u : min(n, gp->n_sorted);
}

__device__ __forceinline__ float frag_depth(const DepthParams& dp, float z_c, float l2)
{
	float const off = cos_half_pi(sqrtr(l2));           // depth.frag:25
	float const zf = subr(z_c, mulr(dp.h, off));        // depth.frag:28
	float d = divr(addr(mulr(dp.P22, zf), dp.P32), zf); // depth.frag:30-33
	return d < 0.0f ? 0.0f : (d > 1.0f ? 1.0f : d);
}

// The same depth to within two ulps, for BOUNDS only (which carry 16 ulps of slack): approximate square root and
// reciprocal, the cosine's polynomial and the rest contracted to FMAs -- 17 instructions instead of 45.
__device__ __forceinline__ float frag_depth_bound(const DepthParams& dp, float z_c, float l2)
{
	float sq;
	asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(l2));
	float const x = 1.57079632679f * sq, x2 = x * x;
	float p = 2.08767569878681e-9f;
	p = fmaf(p, x2, -2.75573192239859e-7f);
	p = fmaf(p, x2, 2.48015873015873e-5f);
	p = fmaf(p, x2, -1.38888888888889e-3f);
	p = fmaf(p, x2, 4.16666666666667e-2f);
	p = fmaf(p, x2, -0.5f);
	p = fmaf(p, x2, 1.0f);
	float const zf = fmaf(-dp.h, p, z_c);
	float const d = fmaf(dp.P32, __fdividef(1.0f, zf), dp.P22);
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
	// (an affine view matrix -- last row 0 0 0 1, every camera of the reference -- makes w exactly 1 and x / 1 = x)
	float const x_c = dp.affine_view ? vc[0] : divr(vc[0], vc[3]), y_c = dp.affine_view ? vc[1] : divr(vc[1], vc[3]),
		z_c = dp.affine_view ? vc[2] : divr(vc[2], vc[3]);
	if (!(z_c > 0.0f)) return false;
	float const quad_depth = divr(addr(mulr(dp.P22, z_c), dp.P32), z_c);
	if (!(quad_depth >= 0.0f && quad_depth <= 1.0f)) return false;   // quad clipped by near / far

	// conservative pixel box of the disc (any superset yields the same image): one approximate reciprocal serves its
	// four quotients -- their error, ~1e-3 of a pixel, disappears in the whole pixel of slack the radii carry
	float inv_z;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_z) : "f"(z_c));
	float const cx = (dp.P00 * x_c * inv_z + 1.0f) * dp.half_w;
	float const cy = (dp.P11 * y_c * inv_z + 1.0f) * dp.half_h;
	float const rx = fabsf(dp.P00) * dp.h * inv_z * dp.half_w + 1.0f;
	float const ry = fabsf(dp.P11) * dp.h * inv_z * dp.half_h + 1.0f;
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

// pass 1: per-particle splat parameters -> scratch, and the cheapest useful bound: the tile that contains the
// disc centre (fully covered whenever the disc radius exceeds the tile diagonal, which is how T is chosen)
template <int T>
__global__ void __launch_bounds__(256) k_depth_seed(const float4* __restrict__ sorted, const float* __restrict__ raw, uint32_t n, const GridParams* __restrict__ gp, DepthParams dp,
													float4* __restrict__ splat_a, uint4* __restrict__ splat_b,
													uint32_t* __restrict__ tile_bound, uint32_t* __restrict__ depth_bits, uint32_t npix)
{
	pdl_enter();
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	// depth <- 1.0 (DepthRenderPass.cpp:54) rides along here: nothing reads the image before k_depth_splat, and the 8 MB of
	// stores overlap the projection arithmetic instead of standing at the head of the chain in k_depth_clear
	{
		uint32_t const rw = dp.rx1 > 0 ? (uint32_t)(dp.rx1 - dp.rx0) : (uint32_t)dp.W;
		for (uint32_t p = i; p < npix; p += gridDim.x * blockDim.x)
			depth_bits[rw == (uint32_t)dp.W ? (size_t)p : (size_t)((uint32_t)dp.ry0 + p / rw) * (uint32_t)dp.W + (uint32_t)dp.rx0 + p % rw] = 0x3f800000u;
	}
	if (i >= sorted_count(gp, n)) return;
	Splat s;
	float4 const p = raw ? make_float4(__ldg(raw + 3ull * i), __ldg(raw + 3ull * i + 1), __ldg(raw + 3ull * i + 2), 0.0f) : __ldg(sorted + i);
	bool live = splat_setup(dp, p, s);
	if (live && !box_owned(dp, s.x0, s.y0, s.x1, s.y1)) live = false;      // cannot touch a pixel this rank renders
	splat_a[i] = make_float4(s.z_c, s.ax, s.bx, s.ay);
	splat_b[i] = make_uint4(__float_as_uint(s.by), s.near_bits, live ? ((uint32_t)s.x0 | ((uint32_t)s.x1 << 16)) : 0xffffu,
							live ? ((uint32_t)s.y0 | ((uint32_t)s.y1 << 16)) : 0xffffu);
	if (!live) return;
	// tile under the disc centre: u = 0 at ndc_x = bx / ax  ->  pixel = (ndc + 1) * W/2 - 0.5 (which tile gets the bound is a
	// choice -- the bound itself is checked against the tile's corners below -- so approximate quotients do)
	float const pcx = (__fdividef(s.bx, s.ax) + 1.0f) * dp.half_w - 0.5f;
	float const pcy = (__fdividef(s.by, s.ay) + 1.0f) * dp.half_h - 0.5f;
	if (!(pcx >= 0.0f && pcy >= 0.0f && pcx < (float)dp.W && pcy < (float)dp.H)) return;
	int const tx = (int)pcx / T, ty = (int)pcy / T;
	int const px0 = tx * T, px1 = min(px0 + T - 1, dp.W - 1);
	int const py0 = ty * T, py1 = min(py0 + T - 1, dp.H - 1);
	float const u0 = frag_u((float)px0, dp.two_w_inv, s.ax, s.bx), u1 = frag_u((float)px1, dp.two_w_inv, s.ax, s.bx);
	float const v0 = frag_u((float)py0, dp.two_h_inv, s.ay, s.by), v1 = frag_u((float)py1, dp.two_h_inv, s.ay, s.by);
	// largest l2 any pixel of the tile evaluates to (FP32 mul/add are monotone)
	float const l2 = addr(fmaxf(mulr(u0, u0), mulr(u1, u1)), fmaxf(mulr(v0, v0), mulr(v1, v1)));
	if (l2 > 1.0f) return;                                // some pixel of the tile is discarded: no bound
	float const d = frag_depth_bound(dp, s.z_c, l2);
	uint32_t const bound = __float_as_uint(d) + 16u;      // slack for the non-monotone last bits of the cosine and frag_depth_bound's two ulps
	if (bound >= 0x3f800000u) return;
	uint32_t* const cell = tile_bound + (size_t)ty * dp.tiles_x + tx;
	if (bound < __ldcg(cell)) atomicMin(cell, bound);
}

__device__ __forceinline__ bool splat_load(const float4* __restrict__ splat_a, const uint4* __restrict__ splat_b,
										   uint32_t i, Splat& s)
{
	float4 const a = __ldg(splat_a + i);
	uint4 const b = __ldg(splat_b + i);
	s.z_c = a.x; s.ax = a.y; s.bx = a.z; s.ay = a.w;
	s.by = __uint_as_float(b.x); s.near_bits = b.y;
	s.x0 = (int)(b.z & 0xffffu); s.x1 = (int)(b.z >> 16);
	s.y0 = (int)(b.w & 0xffffu); s.y1 = (int)(b.w >> 16);
	return s.x1 >= s.x0;
}

// appends the `take` threads of a 256-thread CTA to a list with ONE atomic per CTA (one per warp is thousands of
// atomics on a single address per kernel: they queue up at the L2); returns the thread's slot or 0xffffffff
__device__ __forceinline__ uint32_t cta_append(bool take, uint32_t* __restrict__ counter)
{
	__shared__ uint32_t s_warp_cnt[8];
	__shared__ uint32_t s_cta_base;
	uint32_t const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t const m = __ballot_sync(0xffffffffu, take);
	if (lane == 0) s_warp_cnt[warp] = (uint32_t)__popc(m);
	__syncthreads();
	if (threadIdx.x == 0)
	{
		uint32_t total = 0;
#pragma unroll
		for (int w = 0; w < 8; w++) total += s_warp_cnt[w];
		s_cta_base = total ? atomicAdd(counter, total) : 0u;
	}
	__syncthreads();
	if (!take) return 0xffffffffu;
	uint32_t slot = s_cta_base + (uint32_t)__popc(m & ((1u << lane) - 1u));
	for (uint32_t w = 0; w < warp; w++) slot += s_warp_cnt[w];
	return slot;
}

// pass 2 (optional refinement): every tile the disc covers completely gets the particle's bound.  Only particles
// that are still in front of the seed bound of their own centre tile take part -- the fluid's front layers, which sit
// next to each other in the cell-sorted array: handled by the thread that owns the particle they kept a few CTAs busy
// while the rest of the GPU idled (ncu r01 final: SMs 58 % active, 28 % occupancy).  So k_depth_gate only lists them
// (warp-aggregated append) and k_depth_bounds deals the list out evenly, one warp per particle, lanes = tiles.
template <int T>
__global__ void __launch_bounds__(256) k_depth_gate(uint32_t n, const GridParams* __restrict__ gp, DepthParams dp, const uint4* __restrict__ splat_b,
													const uint32_t* __restrict__ tile_bound, uint32_t* __restrict__ list,
													uint32_t* __restrict__ n_list)
{
	pdl_enter();
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	bool take = false;
	if (i < sorted_count(gp, n))
	{
		uint4 const b = __ldg(splat_b + i);
		int const x0 = (int)(b.z & 0xffffu), x1 = (int)(b.z >> 16), y0 = (int)(b.w & 0xffffu), y1 = (int)(b.w >> 16);
		if (x1 >= x0)
		{
			int const cx = (x0 / T + x1 / T) >> 1, cy = (y0 / T + y1 / T) >> 1;
			take = b.y < __ldcg(tile_bound + (size_t)cy * dp.tiles_x + cx);
		}
	}
	uint32_t const slot = cta_append(take, n_list);
	if (take) list[slot] = i;
}

template <int T>
__global__ void __launch_bounds__(256) k_depth_bounds(DepthParams dp, const float4* __restrict__ splat_a,
													  const uint4* __restrict__ splat_b, const uint32_t* __restrict__ list,
													  const uint32_t* __restrict__ n_list, uint32_t* __restrict__ tile_bound)
{
	pdl_enter();
	int const lane = threadIdx.x & 31;
	uint32_t const nwarps = (gridDim.x * blockDim.x) >> 5;
	uint32_t const count = __ldcg(n_list);
	for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < count; w += nwarps)
	{
		Splat s;
		splat_load(splat_a, splat_b, __ldg(list + w), s);
		int tx0 = s.x0 / T, tx1 = s.x1 / T, ty0 = s.y0 / T, ty1 = s.y1 / T;
		// the pixel box is the disc plus a margin of a pixel, so the tile that holds an unclipped box edge has pixel
		// centres outside the disc and cannot be covered completely: leave the rim out (6 x 6 -> 4 x 4 tiles at C2)
		if (s.x0 > 0) tx0++;
		if (s.x1 < dp.W - 1) tx1--;
		if (s.y0 > 0) ty0++;
		if (s.y1 < dp.H - 1) ty1--;
		if (tx1 < tx0 || ty1 < ty0) continue;
		int const nx = tx1 - tx0 + 1, ntile = nx * (ty1 - ty0 + 1);
		// t / nx by one multiplication: (t + 0.5) / nx is at least 0.5 / nx away from an integer and the product's error
		// below 2^16 * 2^-22 / nx
		float const inv_nx = 1.0f / (float)nx;
		for (int t = lane; t < ntile; t += 32)
		{
			int const row = ntile < (1 << 16) ? (int)(((float)t + 0.5f) * inv_nx) : t / nx;
			int const ty = ty0 + row, tx = tx0 + (t - row * nx);
			int const px0 = tx * T, px1 = min(px0 + T - 1, dp.W - 1);
			int const py0 = ty * T, py1 = min(py0 + T - 1, dp.H - 1);
			if (!pixel_owned(dp, px0, py0)) continue;
			float const v0 = frag_u((float)py0, dp.two_h_inv, s.ay, s.by), v1 = frag_u((float)py1, dp.two_h_inv, s.ay, s.by);
			float const u0 = frag_u((float)px0, dp.two_w_inv, s.ax, s.bx), u1 = frag_u((float)px1, dp.two_w_inv, s.ax, s.bx);
			float const l2 = addr(fmaxf(mulr(u0, u0), mulr(u1, u1)), fmaxf(mulr(v0, v0), mulr(v1, v1)));
			if (l2 > 1.0f) continue;
			float const d = frag_depth_bound(dp, s.z_c, l2);
			uint32_t const bound = __float_as_uint(d) + 16u;
			if (bound >= 0x3f800000u) continue;
			uint32_t* const cell = tile_bound + (size_t)ty * dp.tiles_x + tx;
			if (bound < __ldcg(cell)) atomicMin(cell, bound);
		}
	}
}

// coarse level of the tile bounds: the maximum over every 4 x 4 block of tiles.  A particle whose nearest possible
// depth is not in front of that maximum cannot win any tile of the block.
constexpr int kCoarse = 4;

// One CTA per 16 x 16 block of tiles: thread t holds tile (t % 16, t / 16) of the block; the 4 x 4 maxima go to
// `coarse`, the maximum of the whole block to `coarse2` -- the level that settles a particle deep inside the fluid
// with one to four loads (k_depth_cull).
constexpr int kCoarse2 = 4 * kCoarse;

__global__ void __launch_bounds__(256) k_depth_coarse(const uint32_t* __restrict__ tile_bound, int tiles_x, int tiles_y,
													  uint32_t* __restrict__ coarse, int cx, uint32_t* __restrict__ coarse2, int c2x)
{
	pdl_enter();
	int const bx2 = blockIdx.x % c2x, by2 = blockIdx.x / c2x;
	int const lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
	int const tx = bx2 * kCoarse2 + lx, ty = by2 * kCoarse2 + ly;
	uint32_t m = 0u;
	if (tx < tiles_x && ty < tiles_y) m = __ldg(tile_bound + (size_t)ty * tiles_x + tx);
	// 4 x 4 maxima: over lx within groups of 4 (lanes 0..3 | 4..7 | ...), then over ly within groups of 4 through shared memory
	m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
	m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
	__shared__ uint32_t s_row[16][4];
	__shared__ uint32_t s_blk[16];
	if ((lx & 3) == 0) s_row[ly][lx >> 2] = m;
	__syncthreads();
	if (threadIdx.x < 16)
	{
		int const gx = threadIdx.x & 3, gy = threadIdx.x >> 2;
		uint32_t v = max(max(s_row[4 * gy][gx], s_row[4 * gy + 1][gx]), max(s_row[4 * gy + 2][gx], s_row[4 * gy + 3][gx]));
		int const ccx = bx2 * 4 + gx, ccy = by2 * 4 + gy;
		if (ccx < cx && ccy * kCoarse < tiles_y) coarse[(size_t)ccy * cx + ccx] = v;
		s_blk[threadIdx.x] = v;
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		uint32_t v = 0u;
#pragma unroll
		for (int k = 0; k < 16; k++) v = max(v, s_blk[k]);
		coarse2[blockIdx.x] = v;
	}
}

// pass 3: keep the particles that can still win a pixel of some tile they overlap (compacted, warp-aggregated).
// Interior particles -- almost all of them -- are settled by the few coarse blocks their box touches.
template <int T>
__global__ void __launch_bounds__(256) k_depth_cull(uint32_t n, const GridParams* __restrict__ gp, DepthParams dp, const uint4* __restrict__ splat_b,
													const uint32_t* __restrict__ tile_bound, const uint32_t* __restrict__ coarse,
													int coarse_x, const uint32_t* __restrict__ coarse2, int c2x,
													uint32_t* __restrict__ survivors, uint32_t* __restrict__ n_survivors)
{
	pdl_enter();
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	bool wins = false;
	if (i < sorted_count(gp, n))
	{
		uint4 const b = __ldg(splat_b + i);
		int const x0 = (int)(b.z & 0xffffu), x1 = (int)(b.z >> 16), y0 = (int)(b.w & 0xffffu), y1 = (int)(b.w >> 16);
		if (x1 >= x0)
		{
			int const tx0 = x0 / T, ty0 = y0 / T, tx1 = x1 / T, ty1 = y1 / T;
			// the 16 x 16-tile blocks first: a particle behind the bound of every one it touches (nearly all particles) is done
			bool maybe = false;
			for (int by2 = ty0 / kCoarse2; by2 <= ty1 / kCoarse2; by2++)
				for (int bx2 = tx0 / kCoarse2; bx2 <= tx1 / kCoarse2; bx2++) maybe = maybe || b.y < __ldg(coarse2 + (size_t)by2 * c2x + bx2);
			if (maybe)
			for (int by = ty0 / kCoarse; by <= ty1 / kCoarse && !wins; by++)
				for (int bx = tx0 / kCoarse; bx <= tx1 / kCoarse && !wins; bx++)
				{
					if (!pixel_owned(dp, bx * kCoarse * T, by * kCoarse * T)) continue;
					if (b.y >= __ldg(coarse + (size_t)by * coarse_x + bx)) continue;
					int const fy0 = max(ty0, by * kCoarse), fy1 = min(ty1, by * kCoarse + kCoarse - 1);
					int const fx0 = max(tx0, bx * kCoarse), fx1 = min(tx1, bx * kCoarse + kCoarse - 1);
					for (int ty = fy0; ty <= fy1 && !wins; ty++)
					{
						const uint32_t* row = tile_bound + (size_t)ty * dp.tiles_x;
						for (int tx = fx0; tx <= fx1; tx++)
							if (b.y < __ldg(row + tx)) { wins = true; break; }
					}
				}
		}
	}
	uint32_t const slot = cta_append(wins, n_survivors);
	if (wins) survivors[slot] = i;
}

// pass 4: one warp per surviving particle: lanes re-test the tiles in parallel, then evaluate the fragments of the
// tiles that are still open, 32 pixels at a time
template <int T>
__global__ void __launch_bounds__(256) k_depth_splat(DepthParams dp, const float4* __restrict__ splat_a,
													 const uint4* __restrict__ splat_b,
													 const uint32_t* __restrict__ survivors, const uint32_t* __restrict__ n_survivors,
													 const uint32_t* __restrict__ tile_bound, uint32_t* __restrict__ depth_bits)
{
	pdl_enter();
	constexpr int PIX = T * T;                       // pixels per tile
	constexpr int LPT = PIX < 32 ? PIX : 32;         // lanes working on one tile
	constexpr int TPR = 32 / LPT;                    // tiles per round
	constexpr int RPT = PIX / LPT;                   // rounds per tile
	uint32_t const lane = threadIdx.x & 31u;
	__shared__ int s_open[8][32];
	uint32_t const count = __ldcg(n_survivors);
	// (survivors are dealt out statically: one atomic ticket per warp costs more than the imbalance it removes,
	// 9 472 warps on one counter -- r01q)
	uint32_t const nwarps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < count; w += nwarps)
	{
		Splat s;
		splat_load(splat_a, splat_b, __ldg(survivors + (dp.reverse ? count - 1u - w : w)), s);
		int const ptx0 = s.x0 / T, pty0 = s.y0 / T;
		int const pntx = s.x1 / T - ptx0 + 1;
		int const ntiles = pntx * (s.y1 / T - pty0 + 1);
		float const inv_ntx = 1.0f / (float)pntx;

		for (int tb = 0; tb < ntiles; tb += 32)
		{
			int const t = tb + (int)lane;
			// t / pntx for the small ints involved
			int trow = (int)(((float)t + 0.5f) * inv_ntx);
			int tcol = t - trow * pntx;
			if (tcol < 0) { trow--; tcol += pntx; } else if (tcol >= pntx) { trow++; tcol -= pntx; }
			// a tile stays open only if the NEAREST fragment the particle can produce on it -- the pixel centre closest
			// to the disc centre: u and v are monotone in the pixel index, so per axis that is 0 when the tile straddles
			// the centre line and the smaller end otherwise; FP32 mul / add are monotone, the polynomial cosine is up
			// to a few ulps (16 of slack) -- is in front of the tile's bound.  The disc-centre depth (near_bits) only
			// tells front layers from interior ones: with discs of ~12 px radius at ~5 px spacing a front-layer particle
			// overlaps ~50 tiles but can win only the handful next to its centre.
			bool alive = t < ntiles && pixel_owned(dp, (ptx0 + tcol) * T, (pty0 + trow) * T);
			if (alive)
			{
				int const tx = ptx0 + tcol, ty = pty0 + trow;
				int const qx0 = max(tx * T, s.x0), qx1 = min(tx * T + T - 1, s.x1);
				int const qy0 = max(ty * T, s.y0), qy1 = min(ty * T + T - 1, s.y1);
				float const u0 = frag_u((float)qx0, dp.two_w_inv, s.ax, s.bx), u1 = frag_u((float)qx1, dp.two_w_inv, s.ax, s.bx);
				float const v0 = frag_u((float)qy0, dp.two_h_inv, s.ay, s.by), v1 = frag_u((float)qy1, dp.two_h_inv, s.ay, s.by);
				float const un = (u0 <= 0.0f) != (u1 <= 0.0f) ? 0.0f : fminf(fabsf(u0), fabsf(u1));
				float const vn = (v0 <= 0.0f) != (v1 <= 0.0f) ? 0.0f : fminf(fabsf(v0), fabsf(v1));
				float const l2n = addr(mulr(un, un), mulr(vn, vn));
				alive = l2n <= 1.0f;                                   // else every pixel of the tile is discarded (depth.frag:22)
				if (alive)
				{
					uint32_t const dn = __float_as_uint(frag_depth_bound(dp, s.z_c, l2n));
					alive = (dn > 24u ? dn - 24u : 0u) < __ldg(tile_bound + (size_t)ty * dp.tiles_x + tx);
				}
			}
			uint32_t const tiles = __ballot_sync(0xffffffffu, alive);
			if (tiles == 0u) continue;
			// the open tiles of this batch, compacted into the warp's shared-memory row: lane group g of round k takes
			// entry k * TPR + g (a __fns per round cost 30% of the kernel's instructions, r01 final)
			int const n_open = __popc(tiles);
			__syncwarp();
			if (alive) s_open[threadIdx.x >> 5][__popc(tiles & ((1u << lane) - 1u))] = ((pty0 + trow) << 16) | (ptx0 + tcol);
			__syncwarp();
			for (int ob = 0; ob < n_open; ob += TPR)
			{
				int const o = ob + (int)(lane / LPT);
				if (o >= n_open) continue;
				int const txy = s_open[threadIdx.x >> 5][o];
				int const tpx = (txy & 0xffff) * T, tpy = (txy >> 16) * T;
#pragma unroll
				for (int r = 0; r < RPT; r++)
				{
					int const k = r * LPT + (int)(lane % LPT);
					int const px = tpx + (k % T), py = tpy + (k / T);
					if (px < s.x0 || px > s.x1 || py < s.y0 || py > s.y1) continue;
					uint32_t* const cell = depth_bits + (size_t)py * (size_t)dp.W + (size_t)px;
					uint32_t const cur = __ldcg(cell);                  // L2 read: always fresh; in flight during the arithmetic below
					float const u = frag_u((float)px, dp.two_w_inv, s.ax, s.bx);
					float const v = frag_u((float)py, dp.two_h_inv, s.ay, s.by);
					float const l2 = addr(mulr(u, u), mulr(v, v));
					if (s.near_bits >= cur) continue;                   // cannot win this pixel
					if (l2 > 1.0f) continue;                            // depth.frag:22 `if (l2 > 1) discard;`
					float const d = frag_depth(dp, s.z_c, l2);
					if (!(d < 1.0f)) continue;                          // compare Less against the clear value
					atomicMin(cell, __float_as_uint(d));
				}
			}
		}
	}
}

template <int T>
int launch_tiles(Context* ctx, const Frame& f, DepthParams dp, bool refine_bounds, cudaStream_t st, const float* raw, uint32_t npix)
{
	uint32_t const n = (uint32_t)f.n;
	uint32_t const blocks = (n + 255u) / 256u;
	const GridParams* const gp = raw ? nullptr : f.d_gp;
	float4* const splat_a = (float4*)ctx->d_splat;
	uint4* const splat_b = (uint4*)(ctx->d_splat + 4 * (size_t)n);
	uint32_t* const n_surv = ctx->d_survivors;
	uint32_t* const surv = ctx->d_survivors + 4;
	FM_CUDA(launch_pdl(k_depth_seed<T>, dim3(blocks), dim3(256), 0, st, (const float4*)f.d_sorted, raw, n, gp, dp, splat_a, splat_b, ctx->d_tile_bound,
					   (uint32_t*)ctx->d_depth, npix));
	uint32_t const want = (n + 7u) / 8u;
	uint32_t const cap = (uint32_t)ctx->sm_count * 8u;
	if (refine_bounds)
	{
		// the gate's list shares the survivor array (it is consumed before k_depth_cull writes there); its count is word 2
		FM_CUDA(launch_pdl(k_depth_gate<T>, dim3(blocks), dim3(256), 0, st, n, gp, dp, splat_b, ctx->d_tile_bound, surv, n_surv + 2));
		FM_CUDA(launch_pdl(k_depth_bounds<T>, dim3(want < cap ? want : cap), dim3(256), 0, st, dp, splat_a, splat_b, surv, n_surv + 2, ctx->d_tile_bound));
	}
	int const cx = (dp.tiles_x + kCoarse - 1) / kCoarse, cy = (dp.tiles_y + kCoarse - 1) / kCoarse;
	uint32_t* const coarse = ctx->d_tile_bound + (size_t)dp.tiles_x * dp.tiles_y;
	int const c2x = (dp.tiles_x + kCoarse2 - 1) / kCoarse2, c2y = (dp.tiles_y + kCoarse2 - 1) / kCoarse2;
	uint32_t* const coarse2 = coarse + (size_t)cx * cy;
	FM_CUDA(launch_pdl(k_depth_coarse, dim3(c2x * c2y), dim3(256), 0, st, ctx->d_tile_bound, dp.tiles_x, dp.tiles_y, coarse, cx, coarse2, c2x));
	FM_CUDA(launch_pdl(k_depth_cull<T>, dim3(blocks), dim3(256), 0, st, n, gp, dp, splat_b, ctx->d_tile_bound, coarse, cx, coarse2, c2x, surv, n_surv));
	FM_CUDA(launch_pdl(k_depth_splat<T>, dim3(want < cap ? want : cap), dim3(256), 0, st, dp, splat_a, splat_b, surv, n_surv, ctx->d_tile_bound,
															 (uint32_t*)ctx->d_depth));
	ctx->kernel_launches += refine_bounds ? 6 : 4;
	return FR_OK;
}

}  // namespace

int launch_depth_prepass(Context* ctx, const Frame& f, cudaStream_t st, const float* raw_xyz)
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
	dp.affine_view = (cam.view[3] == 0.0f && cam.view[7] == 0.0f && cam.view[11] == 0.0f && cam.view[15] == 1.0f) ? 1 : 0;
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
	// the pre-pass is partitioned with the march when every hi-Z tile / coarse block (<= 64 px) has one owner (tile
	// sizes that are powers of two >= 64);
	// otherwise every rank computes the whole depth image (same result on the pixels it uses)
	auto pow2 = [](int v) { return v >= 64 && (v & (v - 1)) == 0; };
	bool const part_ok = ctx->part_world > 1 && pow2(ctx->part_tw) && pow2(ctx->part_th);
	dp.part_rank = ctx->part_rank; dp.part_world = part_ok ? ctx->part_world : 1;
	dp.part_sw = 0; dp.part_sh = 0;
	while ((1 << dp.part_sw) < ctx->part_tw) dp.part_sw++;
	while ((1 << dp.part_sh) < ctx->part_th) dp.part_sh++;
	dp.part_tiles_x = (ctx->width + ctx->part_tw - 1) / ctx->part_tw;
	bool const region = ctx->region[2] > ctx->region[0];
	dp.rx0 = region ? ctx->region[0] : 0; dp.ry0 = region ? ctx->region[1] : 0;
	dp.rx1 = region ? ctx->region[2] : 0; dp.ry1 = region ? ctx->region[3] : 0;
	if (region) dp.part_world = 1;

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
	size_t const ncoarse = (size_t)((dp.tiles_x + kCoarse - 1) / kCoarse) * ((dp.tiles_y + kCoarse - 1) / kCoarse);
	size_t const ncoarse2 = (size_t)((dp.tiles_x + kCoarse2 - 1) / kCoarse2) * ((dp.tiles_y + kCoarse2 - 1) / kCoarse2);
	if ((rc = ensure_capacity(&ctx->d_tile_bound, &ctx->cap_tile_bound, (size_t)ntiles + ncoarse + ncoarse2))) return rc;   // tiles + both coarse levels
	if ((rc = ensure_capacity(&ctx->d_splat, &ctx->cap_splat, 8 * f.n))) return rc;            // 2 x 16 B per particle
	if ((rc = ensure_capacity(&ctx->d_survivors, &ctx->cap_survivors, f.n + 4))) return rc;     // [0] = count

	// (under a region partition only the region's pixels are cleared -- and only they are valid afterwards)
	uint32_t const rw = region ? (uint32_t)(dp.rx1 - dp.rx0) : (uint32_t)ctx->width, rh = region ? (uint32_t)(dp.ry1 - dp.ry0) : (uint32_t)ctx->height;
	uint32_t const npix = rw * rh;
	// the image itself is cleared by k_depth_seed's threads (one per particle) unless they are too few for it
	bool const seed_clears = (size_t)f.n * 16u >= (size_t)npix;
	uint32_t const counter_threads = ctx->zero_counters_in_depth ? (uint32_t)(sizeof(DeviceCounters) / 4) : 4u;
	uint32_t const small_threads = ntiles > counter_threads ? ntiles : counter_threads;
	uint32_t const clear_threads = seed_clears ? small_threads : (npix > small_threads ? npix : small_threads);
	uint32_t const npix_seed = seed_clears ? npix : 0u;
	FM_CUDA(launch_pdl(k_depth_clear, dim3((clear_threads + 255) / 256), dim3(256), 0, st, (uint32_t*)ctx->d_depth, seed_clears ? 0u : npix, rw, (uint32_t)dp.rx0, (uint32_t)dp.ry0, (uint32_t)ctx->width,
																		  ctx->d_tile_bound, ntiles, ctx->d_survivors,
																	  (uint32_t*)ctx->d_counters, ctx->zero_counters_in_depth ? (uint32_t)(sizeof(DeviceCounters) / 4) : 0u));
	bool const refine = ctx->depth_refine_bounds;
	switch (T)
	{
	case 2: launch_tiles<2>(ctx, f, dp, refine, st, raw_xyz, npix_seed); break;
	case 4: launch_tiles<4>(ctx, f, dp, refine, st, raw_xyz, npix_seed); break;
	case 8: launch_tiles<8>(ctx, f, dp, refine, st, raw_xyz, npix_seed); break;
	default: launch_tiles<16>(ctx, f, dp, refine, st, raw_xyz, npix_seed); break;
	}
	ctx->kernel_launches += 1;
	FM_CUDA(cudaGetLastError());
	return FR_OK;
}

}  // namespace fm
