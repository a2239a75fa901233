// fm_march.cu -- the ray-march kernels with fused normals and shading (sm_100a).
//
// Replaces RayMarcher::PerPixel_Isotropic (src/app/AdvancedRenderer/RayMarcher.cpp:256-344) together with
// its callees Frame::QueryDensityGrid (src/app/Dataset.cpp:26-47), Dataset::GetNeighbors (:272-280),
// CubicSplineKernel::W/gradW (src/app/Kernel.cpp:16-52), intersectAABB (RayMarcher.cpp:51-62), and the
// fullscreen composition pass (assets/shaders/advanced/composition.frag:37-66,70-122,
// CompositionRenderPass.cpp:313-321), which only ever reads its own pixel and is therefore fused as the
// epilogue of the thread that produced the pixel.
//
// Two kernels (the reference's ThreadPool, src/app/ThreadPool.cpp:38-55, hands out single pixels from one
// atomic counter; most pixels return at once because the depth image says "empty"):
//
//   k_classify   every pixel, uniform work: uncovered pixels (depth == 1) get their zero outputs and the
//                background colour right here; each 8x4 pixel tile with at least one covered pixel is
//                appended to a work list (one atomic per CTA).
//   k_march      persistent warps (a multiple of the SM count) pull tiles off the list through one
//                atomic counter -- the ThreadPool's scheme at warp granularity -- so the expensive rays are
//                spread over all 148 SMs no matter where the fluid sits on screen.  Warp = one 8x4 tile,
//                lanes = rays.  The rays of a tile are ~1 cell apart (pixel footprint at the default camera
//                distance ~ h/10), so the 9 contiguous particle ranges each lane walks are the same
//                addresses across the warp: every LDG.128 of a candidate is a single-sector broadcast out of L1.
//                No neighbour list leaves the SM: the d^2 < h^2 walk collects the in-range candidates of each
//                lane in shared memory and the kernel sums run over that list in the reference's accumulation
//                order (eval_density; fm_common.cuh: FrameView).  The gradient sum of the normal is accumulated
//                together with the density on a ray's first sample (where ~94% of the rays seeded by the depth
//                pre-pass hit).  (An explicit L1 prefetch of the candidate lines in front of the walk cost more
//                issue slots and registers than it saved once the walk loop had no branch in it: r01p.)
//
// This header holds the device code of the march; it is compiled twice: fm_march.cu instantiates the isotropic
// kernels (ANISO = false), fm_aniso.cu -- the one translation unit built with -fmad=false -DFM_NO_FMAD, because the
// anisotropic arithmetic (fm_aniso.cuh) is written as plain C expressions -- instantiates PerPixel_Anisotropic
// (RayMarcher.cpp:346-423, ANISO = true).  Stepping, empty-space skip, queues, hit epilogue and shading are shared.
#pragma once

#include "fm_internal.h"
#ifdef FM_NO_FMAD
#include "fm_aniso.cuh"
#endif

namespace fm
{

// everything the host needs to launch the two march kernels (filled by launch_march)
struct MarchParams
{
	int W, H;
	float two_w_inv, two_h_inv;       // m_TwoWidthInv, m_TwoHeightInv (RayMarcher.cpp:88-89)
	float inv_w, inv_h;
	float ipv[16];                    // m_InvProjectionView
	float cam[3];                     // m_CameraPosition
	float dir[3];                     // Uniforms.CameraDirection
	int max_steps;
	float step_size, iso;
	int bisection_steps;
	int skip_last_pixel;
	int early_out;
	float early_margin;               // cells a sample must lie beyond the grid before the ray is stopped
	int part_rank, part_world, part_tw, part_th, part_tiles_x;
	int rx0, ry0, rx1, ry1;           // region partition: the pixel rectangle this context renders (rx1 == 0: everything)
	int do_march, do_shade;
	int tiles_x;                      // 8x4 pixel tiles per image row
	int aniso_skip_empty;             // anisotropic list evaluation: samples without a particle within h skip the WPCA (FLUIDMARCH_ANISO_EMPTY=0: no)
	int bg_fast;                      // k_classify settles the floor square of an uncovered pixel approximately when it can (FLUIDMARCH_BGFAST=0: never)
	float bg_c[3], bg_err, bg_cam;    // background_fast: ipv[8 + r] + ipv[12 + r], the error bound of its ray, |cam.x| + |cam.z| + 1
	float k_n, k_r, k_s;              // WPCA eigenvalue clamps (RayMarcher.cpp:229-232)
	uint32_t n_eps;
};

struct RayQueues
{
	float4* q1;              // rays left after the first sample -> k_march_long
	uint32_t* ctl;           // [0] tiles listed [1] tile cursor [2] |q1| [3] q1 cursor
	uint32_t* prof;          // FM_LONG_PROFILE builds: four spare words (DeviceCounters::first_examined / first_fallbacks)
};

struct MarchLaunch
{
	FrameView fv;
	MarchParams mp;
	RayQueues rq;
	const uint32_t* tiles;
	uint32_t ctas;
	uint32_t ctas_long;          // k_march_long: CTAs, bitmap words staged in shared memory (0: read from global), dynamic bytes
	uint32_t occ_words;
	size_t smem_long;
	uint32_t occ_words_first;    // k_march_first: the same, only when it costs no resident CTA
	size_t smem_first;
	bool fast_normals;
};

namespace
{

constexpr int kMaxNeighbors = 2 * 4096;   // MAX_NEIGHBORS (RayMarcher.cpp:14)


struct LaneCounters
{
	uint32_t covered, hits, steps, skips, candidates, neighbours, early_exits, overflow;
	uint32_t examined;       // candidates the distance test actually ran on (<= candidates: the staged walk culls cells)
	uint32_t fallbacks;      // samples whose tile did not fit the shared-memory stage (walked out of global memory)
};

#ifndef FM_ANISO_SPEC_WALK
#define FM_ANISO_SPEC_WALK 0              // long_ray's runs of samples in the anisotropic kernel too: measured twice, 2.68 -> 2.90 ms (r03l) and,
                                          // after the empty-neighbourhood skip, 2.02 -> 2.09 ms (r05d): off
#endif
#ifndef FM_LONG_MINBLOCKS
#define FM_LONG_MINBLOCKS 2               // resident 256-thread CTAs per SM k_march_long (isotropic) is compiled for
#endif
#ifndef FM_MARCH_MINBLOCKS
#define FM_MARCH_MINBLOCKS 3              // resident 256-thread CTAs per SM k_march_first is compiled for (<= 85 registers; r02b: 4 CTAs
                                          // at 64 registers and a 32-entry list: C2 0.139 ms, C3 0.450; 3 CTAs with 48 entries: 0.138 / 0.419)
#endif
#ifndef FM_ANISO_MINBLOCKS
#define FM_ANISO_MINBLOCKS 4              // resident CTAs per SM the anisotropic march kernels are compiled for
#endif
#ifndef FM_ANISO_WALK_UNROLL
#define FM_ANISO_WALK_UNROLL 2
#endif
constexpr int kAnisoWalkUnroll = FM_ANISO_WALK_UNROLL;
#ifndef FM_WALK_UNROLL
#define FM_WALK_UNROLL 1
#endif
constexpr int kWalkUnroll = FM_WALK_UNROLL;   // candidates per iteration of the walk loop
#ifndef FM_WALK_FAST
#define FM_WALK_FAST 1                    // ranges that cannot overflow the list are walked by a loop with nothing but the test in it
#endif
#ifndef FM_LIST_CAP
#define FM_LIST_CAP 48                    // in-range candidates a lane collects before it evaluates them
#endif
constexpr int kListCap = FM_LIST_CAP;
constexpr int kListWords = kListCap * 32;  // shared-memory words per warp: list[k * 32 + lane]
// the anisotropic march: one candidate list per warp (aniso_list_build)
#ifndef FM_ANISO_FIRST_LIST
#define FM_ANISO_FIRST_LIST 0             // k_march_first<ANISO> with the warp's list: measured loss (r02o: 1.01 -> 1.30 ms at C2) -- the lanes
                                          // of a tile sit in the same cells, their plain walks already share every load and skip the
                                          // weight arithmetic for the candidates out of everyone's range
#endif
#ifndef FM_ANISO_LIST_CAP
#define FM_ANISO_LIST_CAP 1024
#endif
constexpr uint32_t kAnisoListCap = FM_ANISO_LIST_CAP;                       // entries per warp (u32: sorted index | h-flag << 31)
constexpr size_t kAnisoSmem = (size_t)8 * kAnisoListCap * 4;               // 8 warps per CTA
constexpr size_t kAnisoFirstSmem = FM_ANISO_FIRST_LIST ? kAnisoSmem : 0;   // k_march_first<ANISO> (lists off by default, see below)

// 32-bit shared-window addressing for the per-lane list (a generic uint16_t* is carried as a 64-bit pair with carries)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v)
{
	asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v)
{
	asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
	return v;
}

// ---- the occupancy bitmap into shared memory: one bulk asynchronous copy (TMA, cp.async.bulk) per CTA ------------------
// One thread arms an mbarrier with the byte count and issues the copy; the CTA goes on -- control words, the first tile
// or ray, its depth and step -- and every thread waits on the barrier's phase 0 right before its first look-up.  The
// copy moves the bitmap's 16-byte multiple; the up to three words behind it are stored by plain loads.
__device__ __forceinline__ void bitmap_stage_begin(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint32_t words,
												   unsigned long long* bar)
{
	uint32_t const bar_s = smem_addr(bar), dst_s = smem_addr(dst);
	uint32_t const bulk_words = words & ~3u;
	if (threadIdx.x == 0)
	{
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (threadIdx.x < (words & 3u)) dst[bulk_words + threadIdx.x] = __ldg(src + bulk_words + threadIdx.x);
	__syncthreads();
	if (threadIdx.x == 0)
	{
		if (bulk_words != 0u)
		{
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bulk_words * 4u) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s), "l"(src),
						 "r"(bulk_words * 4u), "r"(bar_s)
						 : "memory");
		}
		else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_s) : "memory");
	}
}

__device__ __forceinline__ void bitmap_stage_wait(unsigned long long* bar)
{
	uint32_t const bar_s = smem_addr(bar);
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"BITMAP_WAIT:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
		"@p bra BITMAP_DONE;\n\t"
		"bra BITMAP_WAIT;\n\t"
		"BITMAP_DONE:\n\t}" ::"r"(bar_s)
		: "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a)
{
	unsigned short v;
	asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
	return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t a)
{
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
	return v;
}

// one neighbour (d = p - x_j, l2 = |d|^2 < h^2) into the running sums, in the reference's list order
template <bool DENS, bool GRAD, bool FAST>
__device__ __forceinline__ void add_neighbour(const FrameView& f, float d0, float d1, float d2, float l2, float& density, f3& g,
											  uint32_t& nn)
{
	if (nn < (uint32_t)kMaxNeighbors)   // list truncation of RayMarcher.cpp:312
	{
		if (DENS && GRAD && !FAST)
		{
			float W;
			f3 gw;
			spline_W_gradW_inrange(f.kernel, mk3(-d0, -d1, -d2), l2, W, gw);
			g = add3(g, gw);
			density = addr(density, W);
			nn++;
			return;
		}
		if (GRAD && !FAST) g = add3(g, spline_gradW_inrange(f.kernel, mk3(-d0, -d1, -d2), l2));
		if (GRAD && FAST)
		{
			float const c = spline_gradW_coeff_fast(f.kernel, l2, mulr(sqrtr(l2), f.kernel.h_inv));
			g.x = fmaf(c, -d0, g.x); g.y = fmaf(c, -d1, g.y); g.z = fmaf(c, -d2, g.z);
		}
		if (DENS) density = addr(density, spline_W_inrange(f.kernel, l2));
	}
	nn++;
}

// density (DENS) and/or the un-normalised gradient sum (GRAD) at p: Dataset::GetNeighbors + the W / gradW
// loops of RayMarcher.cpp:309-336, fused.  Accumulation order == the reference's neighbour order.
//
// Two passes per lane, like the reference's "collect the neighbour list, then sum over it" but without leaving the SM:
// (1) the candidate walk -- load, distance, compare -- appends the indices of the in-range particles to the lane's
// column of a shared-memory list; this loop has no expensive branch in it, so the loads of several candidates are in
// flight at once and all lanes stay active; (2) the kernel sums run over the list.  With the sums inline in the walk
// a warp spent half of its kernel-evaluation slots idle (ncu r01j: 15.4 of 32 lanes active in W / gradW, because
// every lane has its own in-range subset of the shared candidates).  A lane whose list fills up evaluates it and
// resumes the walk where it stopped, so any neighbour count works; the order of the sums is unchanged.
// `list` = this lane's column (stride 32 words); WARP: the whole warp calls together (re-converge between the passes).
template <bool DENS, bool GRAD, bool FAST = false, bool WARP = false>
__device__ __forceinline__ float eval_density(const FrameView& f, f3 p, f3& grad, LaneCounters& lc, uint32_t* __restrict__ list,
											  bool on = true)
{
	int const kx = search_cell_of(f.search_inv, p.x) - f.kmin.x;
	int const ky = search_cell_of(f.search_inv, p.y) - f.kmin.y;
	int const kz = search_cell_of(f.search_inv, p.z) - f.kmin.z;
	int const z0 = max(kz - 1, 0), z1 = min(kz + 1, f.kdim.z - 1);
	float density = 0.0f;
	f3 g = mk3(0.0f, 0.0f, 0.0f);
	uint32_t nn = 0;
	const float4* const sorted = f.sorted;
	float const hh = f.kernel.h_squared;
	bool const walk = z0 <= z1 && on;
	int r0 = 0;              // where the walk resumes after a full list: range r0, particle j0
	uint32_t j0 = 0;
	bool resume = false;
	for (;;)
	{
		uint32_t cnt = 0;
		bool full = false;
		if (walk)
		{
#pragma unroll 1
			for (int r = r0; r < 9 && !full; r++)
			{
				int const x = kx + r / 3 - 1, y = ky + r % 3 - 1;
				if ((unsigned)x >= (unsigned)f.kdim.x || (unsigned)y >= (unsigned)f.kdim.y) continue;
				uint32_t const base = ((uint32_t)x * (uint32_t)f.kdim.y + (uint32_t)y) * (uint32_t)f.kdim.z;
				uint32_t const b = __ldg(f.cell_start + base + z0);
				uint32_t const e = __ldg(f.cell_start + base + z1 + 1);
				uint32_t j = b;
				if (resume && r == r0) j = j0;       // this range was counted when the walk first entered it
				else lc.candidates += e - b;
#if FM_WALK_FAST
				if (cnt + (e - j) <= (uint32_t)kListCap)
				{
					// the list cannot overflow in this range: nothing but the test in the loop, four loads in flight
					uint32_t lp = smem_addr(list) + cnt * 128u;
#define FM_WALK_TEST(q, jj)                                                                          \
	{                                                                                                \
		float const d0 = subr(p.x, (q).x), d1 = subr(p.y, (q).y), d2 = subr(p.z, (q).z);             \
		float const l2 = addr(addr(mulr(d0, d0), mulr(d1, d1)), mulr(d2, d2));                       \
		if (l2 < hh) { sts_u32(lp, (jj)); lp += 128u; }                                              \
	}
#pragma unroll 1
					for (; j + 4u <= e; j += 4u)
					{
						float4 const q0 = __ldg(sorted + j), q1 = __ldg(sorted + j + 1), q2 = __ldg(sorted + j + 2), q3 = __ldg(sorted + j + 3);
						FM_WALK_TEST(q0, j) FM_WALK_TEST(q1, j + 1u) FM_WALK_TEST(q2, j + 2u) FM_WALK_TEST(q3, j + 3u)
					}
#pragma unroll 1
					for (; j < e; j++)
					{
						float4 const q0 = __ldg(sorted + j);
						FM_WALK_TEST(q0, j)
					}
#undef FM_WALK_TEST
					cnt = (lp - smem_addr(list)) >> 7;
					continue;
				}
#endif
				// no branch in the loop: the store and the count are predicated; the first in-range candidate that no
				// longer fits is remembered and the walk resumes there after the list has been evaluated
				uint32_t over = 0xffffffffu;
#pragma unroll (kWalkUnroll)
				for (; j < e; j++)
				{
					float4 const q = __ldg(sorted + j);
					// CompactNSearch: d = x - xb; l2 = d0*d0 + d1*d1 + d2*d2 (left to right); l2 < r2
					float const d0 = subr(p.x, q.x), d1 = subr(p.y, q.y), d2 = subr(p.z, q.z);
					float const l2 = addr(addr(mulr(d0, d0), mulr(d1, d1)), mulr(d2, d2));
					bool const in = l2 < hh;
					bool const fits = cnt < (uint32_t)kListCap;
					if (in && fits) list[cnt * 32u] = j;
					over = (in && !fits) ? min(over, j) : over;
					cnt += (in && fits) ? 1u : 0u;
				}
				if (over != 0xffffffffu) { full = true; resume = true; r0 = r; j0 = over; }
			}
		}
		if (WARP) __syncwarp();
#pragma unroll 1
		for (uint32_t k = 0; k < cnt; k++)
		{
			float4 const q = __ldg(sorted + list[k * 32u]);
			float const d0 = subr(p.x, q.x), d1 = subr(p.y, q.y), d2 = subr(p.z, q.z);
			float const l2 = addr(addr(mulr(d0, d0), mulr(d1, d1)), mulr(d2, d2));
			add_neighbour<DENS, GRAD, FAST>(f, d0, d1, d2, l2, density, g, nn);
		}
		if (WARP)
		{
			if (!__any_sync(0xffffffffu, full)) break;      // the warp stays together while any lane has more to walk
			if (!full) r0 = 9;
		}
		else if (!full) break;
	}
	lc.neighbours += nn;
	if (nn > (uint32_t)kMaxNeighbors) lc.overflow++;
	grad = g;
	return density;
}

// ---- the first sample of a tile out of shared memory ------------------------------------------------------------------
// The 32 rays of an 8x4 pixel tile start within about one search cell of each other, so their 27-cell queries overlap
// almost completely: the union of the cells they touch is a box of 3..4 x 3..4 columns (x, y) times 3..5 cells in z, and
// -- z running fastest in the cell key -- every column of that box is ONE contiguous range of the sorted particle
// array.  The warp copies those ranges into shared memory with coalesced 128-bit loads (one load instruction per 32
// candidates instead of one per candidate and lane), then every lane walks its own 9 sub-ranges out of shared memory:
// LDS.128 broadcasts, no global-memory latency in the inner loop, a loop body of a dozen instructions.  In-range
// candidates are collected as 16-bit stage indices (list of 64 per lane), and the W / gradW sums run over that list
// in the reference's order, reading the particles from the stage as well.  A lane also skips the cells of its 27 that
// its search sphere cannot reach (corner and edge columns: ~24 % of the candidates), with a margin that covers every
// rounding of the cell assignment -- a skipped particle could not have passed `l2 < h^2` (see cull_margins).
// Tiles whose union does not fit (depth discontinuities inside the tile) use the global-memory walk (eval_density).
#ifndef FM_STAGE_CAP
#define FM_STAGE_CAP 512                  // particles a warp can stage (16 B each)
#endif
#ifndef FM_FIRST_WARPS
#define FM_FIRST_WARPS 4                  // warps per CTA of the isotropic k_march_first
#endif
#ifndef FM_FIRST_MINBLOCKS
#define FM_FIRST_MINBLOCKS 4
#endif
#ifndef FM_STAGE_CULL
#define FM_STAGE_CULL 1
#endif
#ifndef FM_STAGE_UNROLL
#define FM_STAGE_UNROLL 4
#endif
constexpr int kStageCap = FM_STAGE_CAP;
constexpr int kStageUnroll = FM_STAGE_UNROLL;
constexpr int kStageCols = 16;            // columns (x, y) of the union box
constexpr int kStageZ = 8;                // cells per column
constexpr int kStageZP = kStageZ + 1;
constexpr int kList16Cap = 64;            // in-range candidates a lane collects before it evaluates them

struct alignas(16) WarpStage
{
	float4 cand[kStageCap];                       // the staged particle ranges, column after column
	uint32_t cs[kStageCols * kStageZP];           // stage offset at which cell k of column c starts (k = nz: its end)
	uint16_t list[kList16Cap * 32];               // list[k * 32 + lane]: stage index of the lane's k-th in-range candidate
};

// Squared lower bounds of |p - x| per axis for particles in the cell below / the own cell / the cell above the sample's
// cell, shrunk by a margin.  A particle stored in search cell c has fl(inv * x) in [c, c + 1] (cell_index truncates the
// rounded product), hence x in [c h, (c + 1) h] up to a relative 2^-22; the face coordinate computed here and the
// subtraction add 2^-23 each.  The margin (|p| + h) * 4e-6 is more than five times their sum, and the threshold the
// bounds are compared with is h^2 (1 + 1e-5), which covers the three roundings of the distance sum itself.
__device__ __forceinline__ void cull_margins(float p, int cell_abs, float h, float& below2, float& above2)
{
	float const face = (float)cell_abs * h;
	float const m = (fabsf(p) + h) * 4e-6f;
	float const lo = fmaxf(p - face - m, 0.0f), hi = fmaxf(face + h - p - m, 0.0f);
	below2 = lo * lo; above2 = hi * hi;
}

// density and (GRAD) gradient sum at p for every lane of the warp; returns false -- for the whole warp, nothing done --
// when the union of the lanes' cells does not fit the stage.  Lanes with on == false take part in the copies only.
template <bool GRAD, bool FAST>
__device__ __forceinline__ bool eval_density_staged(const FrameView& f, f3 p, bool on, WarpStage& ws, float& density_out, f3& grad_out,
													LaneCounters& lc)
{
	constexpr uint32_t FULL = 0xffffffffu;
	int const lane = threadIdx.x & 31;
	// own cell relative to the table, clamped to [-2, kdim + 1] so that +-1 cannot overflow (a lane whose true cell is
	// farther out has an empty range either way)
	int const cx = min(max(search_cell_of(f.search_inv, p.x), f.kmin.x - 2), f.kmin.x + f.kdim.x + 1) - f.kmin.x;
	int const cy = min(max(search_cell_of(f.search_inv, p.y), f.kmin.y - 2), f.kmin.y + f.kdim.y + 1) - f.kmin.y;
	int const cz = min(max(search_cell_of(f.search_inv, p.z), f.kmin.z - 2), f.kmin.z + f.kdim.z + 1) - f.kmin.z;
	int const x0 = max(cx - 1, 0), x1 = min(cx + 1, f.kdim.x - 1);
	int const y0 = max(cy - 1, 0), y1 = min(cy + 1, f.kdim.y - 1);
	int const z0 = max(cz - 1, 0), z1 = min(cz + 1, f.kdim.z - 1);
	bool const valid = on && x0 <= x1 && y0 <= y1 && z0 <= z1;
	int const big = 0x7fffffff;
	int const ux0 = __reduce_min_sync(FULL, valid ? x0 : big), ux1 = __reduce_max_sync(FULL, valid ? x1 : -1);
	density_out = 0.0f;
	grad_out = mk3(0.0f, 0.0f, 0.0f);
	if (ux1 < 0) return true;                               // no lane has anything to walk
	int const uy0 = __reduce_min_sync(FULL, valid ? y0 : big), uy1 = __reduce_max_sync(FULL, valid ? y1 : -1);
	int const uz0 = __reduce_min_sync(FULL, valid ? z0 : big), uz1 = __reduce_max_sync(FULL, valid ? z1 : -1);
	int const ncy = uy1 - uy0 + 1, nz = uz1 - uz0 + 1, ncols = (ux1 - ux0 + 1) * ncy;
	if (ncols > kStageCols || nz > kStageZ) return false;
	int const nzp = nz + 1;
	__syncwarp();                                           // the previous sample's reads of the stage are over
	// cell starts of the union box (global offsets first)
	for (int i = lane; i < ncols * nzp; i += 32)
	{
		int const c = i / nzp, k = i - c * nzp;
		uint32_t const base = ((uint32_t)(ux0 + c / ncy) * (uint32_t)f.kdim.y + (uint32_t)(uy0 + c % ncy)) * (uint32_t)f.kdim.z;
		ws.cs[c * kStageZP + k] = __ldg(f.cell_start + base + (uint32_t)(uz0 + k));
	}
	__syncwarp();
	uint32_t col_b = 0, col_n = 0;                          // lane c < ncols: global start and length of column c
	if (lane < ncols) { col_b = ws.cs[lane * kStageZP]; col_n = ws.cs[lane * kStageZP + nz] - col_b; }
	uint32_t inc = col_n;
#pragma unroll
	for (int o = 1; o < 16; o <<= 1)
	{
		uint32_t const t = __shfl_up_sync(FULL, inc, o);
		if (lane >= o) inc += t;
	}
	uint32_t const col_off = inc - col_n;
	uint32_t const total = __shfl_sync(FULL, inc, 15);
	if (total > (uint32_t)kStageCap) return false;
	if (lane < ncols)
		for (int k = 0; k <= nz; k++) ws.cs[lane * kStageZP + k] += col_off - col_b;      // -> stage offsets
	// the columns, one after the other, 32 particles per load instruction
	for (int c = 0; c < ncols; c++)
	{
		uint32_t const b = __shfl_sync(FULL, col_b, c), n = __shfl_sync(FULL, col_n, c), o = __shfl_sync(FULL, col_off, c);
		for (uint32_t i = lane; i < n; i += 32) ws.cand[o + i] = __ldg(f.sorted + b + i);
	}
	__syncwarp();

	float ax_lo = 0.0f, ax_hi = 0.0f, ay_lo = 0.0f, ay_hi = 0.0f, az_lo = 0.0f, az_hi = 0.0f;
	float const hh_m = f.kernel.h_squared * (1.0f + 1e-5f);
	if (FM_STAGE_CULL)
	{
		cull_margins(p.x, cx + f.kmin.x, f.kernel.h, ax_lo, ax_hi);
		cull_margins(p.y, cy + f.kmin.y, f.kernel.h, ay_lo, ay_hi);
		cull_margins(p.z, cz + f.kmin.z, f.kernel.h, az_lo, az_hi);
	}
	float density = 0.0f;
	f3 g = mk3(0.0f, 0.0f, 0.0f);
	uint32_t nn = 0;
	int r = 0;                       // next of the lane's 9 columns
	int resume_a = -1;               // byte offset into the stage at which a column is resumed after a full list
	uint32_t const my_list = smem_addr(ws.list + lane);                 // shared-window byte addresses
	uint32_t const list_end = my_list + kList16Cap * 64u;
	uint32_t const cb = smem_addr(ws.cand);
	float const hh = f.kernel.h_squared;
	// one candidate at byte offset `off` of the stage.  CompactNSearch: d = x - xb; l2 = d0*d0 + d1*d1 + d2*d2 (left to
	// right); l2 < r2.  The list keeps the byte offset (16 bits: the stage holds at most 4096 particles).
#define FM_STAGE_TEST(q, off)                                                                        \
	{                                                                                                \
		float const d0 = subr(p.x, (q).x), d1 = subr(p.y, (q).y), d2 = subr(p.z, (q).z);             \
		float const l2 = addr(addr(mulr(d0, d0), mulr(d1, d1)), mulr(d2, d2));                       \
		if (l2 < hh) { sts_u16(lp, (off)); lp += 64u; }                                              \
	}
	for (;;)
	{
		uint32_t lp = my_list;           // next free slot of the lane's list
		bool full = false;
		if (valid)
		{
#pragma unroll 1
			while (r < 9 && !full)
			{
				int const dx = r / 3 - 1, dy = r % 3 - 1;
				int const x = cx + dx, y = cy + dy;
				if ((unsigned)x >= (unsigned)f.kdim.x || (unsigned)y >= (unsigned)f.kdim.y) { r++; continue; }
				const uint32_t* const ccs = ws.cs + ((x - ux0) * ncy + (y - uy0)) * kStageZP - uz0;
				int zl = z0, zh = z1;
				if (resume_a < 0) lc.candidates += ccs[z1 + 1] - ccs[z0];
				if (FM_STAGE_CULL)
				{
					float const m2 = (dx == 0 ? 0.0f : (dx < 0 ? ax_lo : ax_hi)) + (dy == 0 ? 0.0f : (dy < 0 ? ay_lo : ay_hi));
					if (m2 >= hh_m) { r++; continue; }
					float const rz2 = hh_m - m2;
					zl = max(az_lo >= rz2 ? cz : cz - 1, 0);
					zh = min(az_hi >= rz2 ? cz : cz + 1, f.kdim.z - 1);
					if (zl > zh) { r++; continue; }
				}
				uint32_t a = ccs[zl] * 16u;
				uint32_t const ae = ccs[zh + 1] * 16u;
				if (resume_a >= 0) { a = (uint32_t)resume_a; resume_a = -1; }
				else lc.examined += (ae - a) >> 4;
				if ((ae - a) >> 4 <= (list_end - lp) >> 6)
				{
					// the list cannot overflow in this column: nothing but the test in the loop, four loads in flight
#pragma unroll 1
					for (; a + 64u <= ae; a += 64u)
					{
						float4 const q0 = lds_f4(cb + a), q1 = lds_f4(cb + a + 16u), q2 = lds_f4(cb + a + 32u), q3 = lds_f4(cb + a + 48u);
						FM_STAGE_TEST(q0, a) FM_STAGE_TEST(q1, a + 16u) FM_STAGE_TEST(q2, a + 32u) FM_STAGE_TEST(q3, a + 48u)
					}
#pragma unroll 1
					for (; a < ae; a += 16u)
					{
						float4 const q0 = lds_f4(cb + a);
						FM_STAGE_TEST(q0, a)
					}
					r++;
				}
				else
				{
#pragma unroll 1
					for (; a < ae; a += 16u)
					{
						float4 const q = lds_f4(cb + a);
						float const d0 = subr(p.x, q.x), d1 = subr(p.y, q.y), d2 = subr(p.z, q.z);
						float const l2 = addr(addr(mulr(d0, d0), mulr(d1, d1)), mulr(d2, d2));
						if (l2 < hh)
						{
							if (lp == list_end) { full = true; resume_a = (int)a; break; }
							sts_u16(lp, a);
							lp += 64u;
						}
					}
					if (!full) r++;
				}
			}
		}
		__syncwarp();
#pragma unroll 1
		for (uint32_t it = my_list; it != lp; it += 64u)
		{
			float4 const q = lds_f4(cb + lds_u16(it));
			float const d0 = subr(p.x, q.x), d1 = subr(p.y, q.y), d2 = subr(p.z, q.z);
			float const l2 = addr(addr(mulr(d0, d0), mulr(d1, d1)), mulr(d2, d2));
			add_neighbour<true, GRAD, FAST>(f, d0, d1, d2, l2, density, g, nn);
		}
		if (!__any_sync(FULL, full)) break;                 // the warp stays together while any lane has more to walk
	}
#undef FM_STAGE_TEST
	lc.neighbours += nn;
	if (nn > (uint32_t)kMaxNeighbors) lc.overflow++;
	density_out = density;
	grad_out = g;
	return true;
}

#ifdef FM_NO_FMAD
// ---- the anisotropic sample: Dataset::GetNeighborsExt + WPCA + AnisotropicKernel sums (RayMarcher.cpp:380-403) -------
// The reference materialises the 2h neighbour list (positions), its h-subset and the WPCA weights in 8192-entry thread
// locals.  Here nothing is materialised: the 27 cells of the r = h_ext search are walked three times in the
// reference's list order -- (1) weights, weight sum, weighted mean; (2) covariance about that mean; (3) kernel sum over
// the h-subset with the finished G -- each pass recomputing membership and weight with the same operations, hence the
// same bits.  Pass 3 (and the gradient pass of a hit) only visits the cells that can hold a particle within h.
struct CellBox { int x0, x1, y0, y1, z0, z1; };

// cells (relative to kmin_ext, clamped to the table and to the query's own 3x3x3 block) that can hold a particle
// within `reach` of p on every axis: the cell index is monotone in the coordinate
__device__ __forceinline__ CellBox ext_box(const FrameView& f, f3 p, float reach)
{
	CellBox b;
	int const cx = search_cell_of(f.search_inv_ext, p.x), cy = search_cell_of(f.search_inv_ext, p.y),
		cz = search_cell_of(f.search_inv_ext, p.z);
	b.x0 = max(max(search_cell_of(f.search_inv_ext, p.x - reach), cx - 1) - f.kmin_ext.x, 0);
	b.x1 = min(min(search_cell_of(f.search_inv_ext, p.x + reach), cx + 1) - f.kmin_ext.x, f.kdim_ext.x - 1);
	b.y0 = max(max(search_cell_of(f.search_inv_ext, p.y - reach), cy - 1) - f.kmin_ext.y, 0);
	b.y1 = min(min(search_cell_of(f.search_inv_ext, p.y + reach), cy + 1) - f.kmin_ext.y, f.kdim_ext.y - 1);
	b.z0 = max(max(search_cell_of(f.search_inv_ext, p.z - reach), cz - 1) - f.kmin_ext.z, 0);
	b.z1 = min(min(search_cell_of(f.search_inv_ext, p.z + reach), cz + 1) - f.kmin_ext.z, f.kdim_ext.z - 1);
	return b;
}

// the query's whole 3x3x3 block (NeighborhoodSearch::find_neighbors)
__device__ __forceinline__ CellBox ext_box_full(const FrameView& f, f3 p)
{
	CellBox b;
	int const cx = search_cell_of(f.search_inv_ext, p.x) - f.kmin_ext.x, cy = search_cell_of(f.search_inv_ext, p.y) - f.kmin_ext.y,
		cz = search_cell_of(f.search_inv_ext, p.z) - f.kmin_ext.z;
	// a sample far outside the table must not overflow cx +- 1
	int const big = 1 << 29;
	int const sx = min(max(cx, -big), big), sy = min(max(cy, -big), big), sz = min(max(cz, -big), big);
	b.x0 = max(sx - 1, 0); b.x1 = min(sx + 1, f.kdim_ext.x - 1);
	b.y0 = max(sy - 1, 0); b.y1 = min(sy + 1, f.kdim_ext.y - 1);
	b.z0 = max(sz - 1, 0); b.z1 = min(sz + 1, f.kdim_ext.z - 1);
	return b;
}

// visits the particles of the box in the reference's result order (x, then y, then z cells; ascending id inside a
// cell); returns the number of particles visited
template <typename Visit>
__device__ __forceinline__ uint32_t walk_ext(const FrameView& f, const CellBox& b, Visit&& visit)
{
	uint32_t visited = 0;
	if (b.z0 > b.z1) return 0;
#pragma unroll 1
	for (int x = b.x0; x <= b.x1; x++)
	{
#pragma unroll 1
		for (int y = b.y0; y <= b.y1; y++)
		{
			uint32_t const base = ((uint32_t)x * (uint32_t)f.kdim_ext.y + (uint32_t)y) * (uint32_t)f.kdim_ext.z;
			uint32_t const jb = __ldg(f.cell_start_ext + base + b.z0);
			uint32_t const je = __ldg(f.cell_start_ext + base + b.z1 + 1);
			visited += je - jb;
#pragma unroll (kAnisoWalkUnroll)
			for (uint32_t j = jb; j < je; j++) visit(__ldg(f.sorted_ext + j));
		}
	}
	return visited;
}

struct AnisoSample
{
	aniso::Mat3 G;
	float detG;
};

__device__ __forceinline__ aniso::Kernel aniso_kernel_of(const FrameView& f)
{
	aniso::Kernel k;
	k.h = f.kernel.h; k.h_squared = f.kernel.h_squared; k.h_inv = f.kernel.h_inv; k.sig = f.aniso_sig;
	return k;
}

// G and det G at p (RayMarcher::WPCA over GetNeighborsExt(p), RayMarcher.cpp:380-400); returns the number of 2h neighbours
__device__ __forceinline__ uint32_t aniso_wpca(const FrameView& f, const MarchParams& mp, f3 p, AnisoSample& as, LaneCounters& lc)
{
	CellBox const full = ext_box_full(f, p);
	float wsum = 0.0f;
	f3 mean = mk3(0.0f, 0.0f, 0.0f);
	uint32_t n_ext = 0;
	lc.candidates += walk_ext(f, full, [&](float4 q) {
		// CompactNSearch: d = x - xb; l2 = d0*d0 + d1*d1 + d2*d2; l2 < r2.  CubicKernel::W takes dot(r, r) of
		// r = xb - x, which is the same sum of the same squares
		float const d0 = p.x - q.x, d1 = p.y - q.y, d2 = p.z - q.z;
		float const l2 = d0 * d0 + d1 * d1 + d2 * d2;
		if (l2 < f.h_ext_squared)
		{
			if (n_ext < (uint32_t)kMaxNeighbors)
			{
				float const w = aniso::cubic_W(f.h_ext, f.search_inv_ext, l2);
				wsum += w;
				mean.x += w * q.x; mean.y += w * q.y; mean.z += w * q.z;
			}
			n_ext++;
		}
	});
	if (n_ext > (uint32_t)kMaxNeighbors) { lc.overflow++; n_ext = (uint32_t)kMaxNeighbors; }
	float const inv_wsum = 1.0f / wsum;
	mean.x *= inv_wsum; mean.y *= inv_wsum; mean.z *= inv_wsum;
	aniso::Sym3 C = { 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
	uint32_t n2 = 0;
	walk_ext(f, full, [&](float4 q) {
		float const d0 = p.x - q.x, d1 = p.y - q.y, d2 = p.z - q.z;
		float const l2 = d0 * d0 + d1 * d1 + d2 * d2;
		if (l2 < f.h_ext_squared)
		{
			if (n2 < (uint32_t)kMaxNeighbors)
			{
				float const w = aniso::cubic_W(f.h_ext, f.search_inv_ext, l2);
				float const x0 = q.x - mean.x, x1 = q.y - mean.y, x2 = q.z - mean.z;
				float const w0 = w * x0, w1 = w * x1, w2 = w * x2;
				C.m00 += w0 * x0;
				C.m10 += w1 * x0; C.m11 += w1 * x1;
				C.m20 += w2 * x0; C.m21 += w2 * x1; C.m22 += w2 * x2;
			}
			n2++;
		}
	});
	C.m00 *= inv_wsum; C.m10 *= inv_wsum; C.m11 *= inv_wsum; C.m20 *= inv_wsum; C.m21 *= inv_wsum; C.m22 *= inv_wsum;
	aniso::Settings st;
	st.k_n = mp.k_n; st.k_r = mp.k_r; st.k_s = mp.k_s; st.n_eps = mp.n_eps;
	aniso::wpca_G(C, n_ext, st, f.kernel.h_inv, as.G);
	as.detG = aniso::det3(as.G);
	return n_ext;
}

// density at p (RayMarcher.cpp:380-403); leaves G, det G in `as` for the hit epilogue
__device__ __forceinline__ float aniso_density(const FrameView& f, const MarchParams& mp, f3 p, AnisoSample& as, LaneCounters& lc)
{
	uint32_t const n_ext = aniso_wpca(f, mp, p, as, lc);
	if (n_ext == 0) return 0.0f;      // no 2h neighbour, hence no h neighbour: the kernel sum is empty
	aniso::Kernel const ak = aniso_kernel_of(f);
	CellBox const near = ext_box(f, p, f.kernel.h * 1.01f);
	float density = 0.0f;
	uint32_t nn = 0;
	walk_ext(f, near, [&](float4 q) {
		f3 const r = mk3(q.x - p.x, q.y - p.y, q.z - p.z);
		float const rr = (r.x * r.x + r.y * r.y) + r.z * r.z;           // glm::dot(r, r) (RayMarcher.cpp:392)
		if (rr < f.kernel.h_squared)
		{
			density += aniso::W(ak, as.G, as.detG, r);
			nn++;
		}
	});
	lc.neighbours += nn;
	return density;
}

// un-normalised normal of a hit at p (RayMarcher.cpp:411-414)
__device__ __forceinline__ f3 aniso_gradient(const FrameView& f, f3 p, const AnisoSample& as)
{
	aniso::Kernel const ak = aniso_kernel_of(f);
	CellBox const near = ext_box(f, p, f.kernel.h * 1.01f);
	f3 g = mk3(0.0f, 0.0f, 0.0f);
	walk_ext(f, near, [&](float4 q) {
		f3 const r = mk3(q.x - p.x, q.y - p.y, q.z - p.z);
		float const rr = (r.x * r.x + r.y * r.y) + r.z * r.z;
		if (rr < f.kernel.h_squared)
		{
			f3 const t = aniso::gradW(ak, as.G, as.detG, r);
			g.x += t.x; g.y += t.y; g.z += t.z;
		}
	});
	return g;
}

// ---- the anisotropic sample with a candidate list shared by the warp ---------------------------------------------------
// The 32 samples a warp evaluates together are close to each other -- the first samples of an 8x4 pixel tile
// (k_march_first), or 32 consecutive samples of one ray (k_march_long) -- while a query's 27 cells of size h_ext hold
// 6-7x the particles within h_ext of it: walking them per lane tests every candidate 32 times, and since nearly every
// candidate is in range of SOME lane the weight / covariance arithmetic runs for all of them at ~15 % lane utilisation.
// Here the warp first lists, lanes = candidates, the particles of the union of the lanes' cell blocks that lie within
// h_ext (+ margin) of the segment the samples sit on (a capsule; for a tile the segment is its diagonal), in
// ascending order of their position in the sorted array.  That order restricted to one lane's 27 cells IS the
// reference's list order (ascending cell key, ascending index inside a cell), and the in-range test each lane then
// applies to every list entry is the reference's own, so sums and truncation see the same particles in the same order:
// bit-identical.  Two things the reference's cell lookup decides rather than the distance are guarded: a particle
// whose distance test passes within rounding of h_ext on one axis could sit two cells away (outside the 27) -- such a
// sample is re-evaluated by the plain walk (`redo`); and a list that does not fit sends the whole warp there.
constexpr int kAnisoListMaxColumns = 400;                                  // larger unions (a window across a long gap) walk plainly

struct SegmentFilter
{
	f3 a, v;
	float inv_vv;
	__device__ __forceinline__ float dist2(float x, float y, float z) const
	{
		float const wx = x - a.x, wy = y - a.y, wz = z - a.z;
		float t = (wx * v.x + wy * v.y + wz * v.z) * inv_vv;
		t = fminf(fmaxf(t, 0.0f), 1.0f);
		float const cx = wx - t * v.x, cy = wy - t * v.y, cz = wz - t * v.z;
		return cx * cx + cy * cy + cz * cz;
	}
};

// returns false when the warp has to fall back to the plain walk; n = entries listed
__device__ __forceinline__ bool aniso_list_build(const FrameView& f, f3 p, bool on, uint32_t* __restrict__ e, uint32_t& n)
{
	uint32_t const FULL = 0xffffffffu;
	uint32_t const lane = threadIdx.x & 31u;
	n = 0;
	if (!(f.h_ext >= 1.02f * f.kernel.h)) return false;                    // the h-subset must lie well inside the list's reach
	CellBox const bx = ext_box_full(f, p);
	bool const has = on && bx.x0 <= bx.x1 && bx.y0 <= bx.y1 && bx.z0 <= bx.z1;
	uint32_t const m_on = __ballot_sync(FULL, has);
	if (m_on == 0u) return true;                                           // nothing in reach of any sample
	int const big = 0x7fffffff;
	int const ux0 = __reduce_min_sync(FULL, has ? bx.x0 : big), ux1 = __reduce_max_sync(FULL, has ? bx.x1 : -big);
	int const uy0 = __reduce_min_sync(FULL, has ? bx.y0 : big), uy1 = __reduce_max_sync(FULL, has ? bx.y1 : -big);
	int const uz0 = __reduce_min_sync(FULL, has ? bx.z0 : big), uz1 = __reduce_max_sync(FULL, has ? bx.z1 : -big);
	if ((ux1 - ux0 + 1) * (uy1 - uy0 + 1) > kAnisoListMaxColumns) return false;
	int const la = __ffs(m_on) - 1, lb = 31 - __clz(m_on);
	SegmentFilter sf;
	sf.a = mk3(__shfl_sync(FULL, p.x, la), __shfl_sync(FULL, p.y, la), __shfl_sync(FULL, p.z, la));
	f3 const b = mk3(__shfl_sync(FULL, p.x, lb), __shfl_sync(FULL, p.y, lb), __shfl_sync(FULL, p.z, lb));
	sf.v = mk3(b.x - sf.a.x, b.y - sf.a.y, b.z - sf.a.z);
	float const vv = sf.v.x * sf.v.x + sf.v.y * sf.v.y + sf.v.z * sf.v.z;
	sf.inv_vv = vv > 0.0f ? 1.0f / vv : 0.0f;
	// how far the samples themselves are from the segment (a ray's samples: rounding only; a tile: half its height)
	float const own = has ? sqrtf(sf.dist2(p.x, p.y, p.z)) : 0.0f;
	float const rho = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(own)));
	float const tol = 1e-3f * f.h_ext + 4e-6f * (fabsf(sf.a.x) + fabsf(sf.a.y) + fabsf(sf.a.z) + fabsf(b.x) + fabsf(b.y) + fabsf(b.z) + 1.0f);
	float const R = f.h_ext + rho + tol, Rh = f.kernel.h + rho + tol;
	float const R2 = R * R * 1.0001f, Rh2 = Rh * Rh * 1.0001f;
	uint32_t const lt = (1u << lane) - 1u;
#pragma unroll 1
	for (int x = ux0; x <= ux1; x++)
	{
#pragma unroll 1
		for (int y = uy0; y <= uy1; y++)
		{
			uint32_t const base = ((uint32_t)x * (uint32_t)f.kdim_ext.y + (uint32_t)y) * (uint32_t)f.kdim_ext.z;
			uint32_t const jb = __ldg(f.cell_start_ext + base + uz0);
			uint32_t const je = __ldg(f.cell_start_ext + base + uz1 + 1);
#pragma unroll 1
			for (uint32_t j0 = jb; j0 < je; j0 += 32u)
			{
				uint32_t const j = j0 + lane;
				bool keep = false;
				uint32_t hflag = 0u;
				if (j < je)
				{
					float4 const q = __ldg(f.sorted_ext + j);
					float const d2 = sf.dist2(q.x, q.y, q.z);
					keep = d2 <= R2;
					hflag = d2 <= Rh2 ? 0x80000000u : 0u;
				}
				uint32_t const bal = __ballot_sync(FULL, keep);
				if (keep)
				{
					uint32_t const pos = n + (uint32_t)__popc(bal & lt);
					if (pos < kAnisoListCap) e[pos] = j | hflag;
				}
				n += (uint32_t)__popc(bal);
			}
		}
	}
	__syncwarp();
	return n <= kAnisoListCap;
}

// the number of particles in the query's 27 cells (what walk_ext reports as visited)
__device__ __forceinline__ uint32_t ext_box_count(const FrameView& f, const CellBox& b)
{
	uint32_t visited = 0;
	if (b.z0 > b.z1) return 0;
	for (int x = b.x0; x <= b.x1; x++)
		for (int y = b.y0; y <= b.y1; y++)
		{
			uint32_t const base = ((uint32_t)x * (uint32_t)f.kdim_ext.y + (uint32_t)y) * (uint32_t)f.kdim_ext.z;
			visited += __ldg(f.cell_start_ext + base + b.z1 + 1) - __ldg(f.cell_start_ext + base + b.z0);
		}
	return visited;
}

// aniso_density over the warp's list: same three passes, same operations.  All lanes run the loops (uniform trip
// count); `on` gates the arithmetic.  `redo` = this lane's sample must be re-evaluated by the plain walk.
__device__ __forceinline__ float aniso_density_list(const FrameView& f, const MarchParams& mp, f3 p, bool on, AnisoSample& as,
													LaneCounters& lc, const uint32_t* __restrict__ e, uint32_t n, bool& redo)
{
	// A sample without a particle within h has density exactly 0 -- the empty sum of RayMarcher.cpp:389-395 -- whatever G is:
	// such a lane skips the weights, the covariance and the eigen-solve (`on` goes false), and a warp of such lanes (a window
	// of a silhouette ray through the fringe of the fluid) skips the passes altogether.  (The 2h-list's overflow counter is
	// not kept for those lanes.)
	if (mp.aniso_skip_empty)
	{
		bool any_h = false;
#pragma unroll 1
		for (uint32_t i = 0; i < n; i++)
		{
			uint32_t const ei = e[i];
			if (!(ei >> 31)) continue;
			float4 const q = __ldg(f.sorted_ext + (ei & 0x7fffffffu));
			f3 const r = mk3(q.x - p.x, q.y - p.y, q.z - p.z);
			float const rr = (r.x * r.x + r.y * r.y) + r.z * r.z;
			any_h = any_h || rr < f.kernel.h_squared;
		}
		if (on && !any_h)
		{
			lc.candidates += ext_box_count(f, ext_box_full(f, p));
			on = false;
		}
		if (!__any_sync(0xffffffffu, on)) { redo = false; return 0.0f; }
	}
	float const edge = f.h_ext * 0.999f - 8e-6f * (fabsf(p.x) + fabsf(p.y) + fabsf(p.z) + f.h_ext);
	float wsum = 0.0f;
	f3 mean = mk3(0.0f, 0.0f, 0.0f);
	uint32_t n_ext = 0;
	redo = false;
#pragma unroll (kAnisoWalkUnroll)
	for (uint32_t i = 0; i < n; i++)
	{
		float4 const q = __ldg(f.sorted_ext + (e[i] & 0x7fffffffu));
		float const d0 = p.x - q.x, d1 = p.y - q.y, d2 = p.z - q.z;
		float const l2 = d0 * d0 + d1 * d1 + d2 * d2;
		if (on && l2 < f.h_ext_squared)
		{
			if (fmaxf(fmaxf(fabsf(d0), fabsf(d1)), fabsf(d2)) > edge) redo = true;
			if (n_ext < (uint32_t)kMaxNeighbors)
			{
				float const w = aniso::cubic_W(f.h_ext, f.search_inv_ext, l2);
				wsum += w;
				mean.x += w * q.x; mean.y += w * q.y; mean.z += w * q.z;
			}
			n_ext++;
		}
	}
	bool const over = n_ext > (uint32_t)kMaxNeighbors;
	if (over) n_ext = (uint32_t)kMaxNeighbors;
	float const inv_wsum = 1.0f / wsum;
	mean.x *= inv_wsum; mean.y *= inv_wsum; mean.z *= inv_wsum;
	aniso::Sym3 C = { 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
	uint32_t n2 = 0;
#pragma unroll (kAnisoWalkUnroll)
	for (uint32_t i = 0; i < n; i++)
	{
		float4 const q = __ldg(f.sorted_ext + (e[i] & 0x7fffffffu));
		float const d0 = p.x - q.x, d1 = p.y - q.y, d2 = p.z - q.z;
		float const l2 = d0 * d0 + d1 * d1 + d2 * d2;
		if (on && l2 < f.h_ext_squared)
		{
			if (n2 < (uint32_t)kMaxNeighbors)
			{
				float const w = aniso::cubic_W(f.h_ext, f.search_inv_ext, l2);
				float const x0 = q.x - mean.x, x1 = q.y - mean.y, x2 = q.z - mean.z;
				float const w0 = w * x0, w1 = w * x1, w2 = w * x2;
				C.m00 += w0 * x0;
				C.m10 += w1 * x0; C.m11 += w1 * x1;
				C.m20 += w2 * x0; C.m21 += w2 * x1; C.m22 += w2 * x2;
			}
			n2++;
		}
	}
	C.m00 *= inv_wsum; C.m10 *= inv_wsum; C.m11 *= inv_wsum; C.m20 *= inv_wsum; C.m21 *= inv_wsum; C.m22 *= inv_wsum;
	aniso::Settings st;
	st.k_n = mp.k_n; st.k_r = mp.k_r; st.k_s = mp.k_s; st.n_eps = mp.n_eps;
	aniso::wpca_G(C, n_ext, st, f.kernel.h_inv, as.G);
	as.detG = aniso::det3(as.G);
	aniso::Kernel const ak = aniso_kernel_of(f);
	bool const on3 = on && n_ext != 0u;        // no 2h neighbour, hence no h neighbour: the kernel sum is empty
	float density = 0.0f;
	uint32_t nn = 0;
#pragma unroll 1
	for (uint32_t i = 0; i < n; i++)
	{
		uint32_t const ei = e[i];
		if (!(ei >> 31)) continue;                                           // (uniform) not within h of any sample of the warp
		float4 const q = __ldg(f.sorted_ext + (ei & 0x7fffffffu));
		f3 const r = mk3(q.x - p.x, q.y - p.y, q.z - p.z);
		float const rr = (r.x * r.x + r.y * r.y) + r.z * r.z;           // glm::dot(r, r) (RayMarcher.cpp:392)
		if (on3 && rr < f.kernel.h_squared)
		{
			density += aniso::W(ak, as.G, as.detG, r);
			nn++;
		}
	}
	if (on && !redo)
	{
		lc.candidates += ext_box_count(f, ext_box_full(f, p));
		lc.neighbours += nn;
		if (over) lc.overflow++;
	}
	return on ? density : 0.0f;
}

// aniso_gradient over the same list, for the lanes whose sample is a hit
__device__ __forceinline__ f3 aniso_gradient_list(const FrameView& f, f3 p, bool on, const AnisoSample& as, const uint32_t* __restrict__ e,
												  uint32_t n)
{
	aniso::Kernel const ak = aniso_kernel_of(f);
	f3 g = mk3(0.0f, 0.0f, 0.0f);
#pragma unroll 1
	for (uint32_t i = 0; i < n; i++)
	{
		uint32_t const ei = e[i];
		if (!(ei >> 31)) continue;
		float4 const q = __ldg(f.sorted_ext + (ei & 0x7fffffffu));
		f3 const r = mk3(q.x - p.x, q.y - p.y, q.z - p.z);
		float const rr = (r.x * r.x + r.y * r.y) + r.z * r.z;
		if (on && rr < f.kernel.h_squared)
		{
			f3 const t = aniso::gradW(ak, as.G, as.detG, r);
			g.x += t.x; g.y += t.y; g.z += t.z;
		}
	}
	return g;
}

// the warp's 32 samples (lanes with `on`): density, and G / det G left in `as`.  listed = the list in `e` (n entries) is
// valid for these samples afterwards (the gradient pass of the hits can use it).
__device__ __forceinline__ float aniso_density_warp(const FrameView& f, const MarchParams& mp, f3 p, bool on, AnisoSample& as,
													LaneCounters& lc, uint32_t* __restrict__ e, uint32_t& n, bool& listed)
{
	__syncwarp();
	listed = aniso_list_build(f, p, on, e, n);
	if (!listed) return on ? aniso_density(f, mp, p, as, lc) : 0.0f;
	bool redo;
	float density = aniso_density_list(f, mp, p, on, as, lc, e, n, redo);
	if (redo) density = aniso_density(f, mp, p, as, lc);
	return density;
}
#endif   // FM_NO_FMAD

// ---- one sample of either flavour ------------------------------------------------------------------------------------
// what a density evaluation leaves behind for the hit epilogue
template <bool ANISO> struct SampleState;
template <> struct SampleState<false> { f3 grad; bool have_grad; };
#ifdef FM_NO_FMAD
template <> struct SampleState<true> { AnisoSample as; f3 grad; bool have_grad; };
#endif

// WITH_GRAD (isotropic only): accumulate the gradient sum together with the density
template <bool ANISO, bool WITH_GRAD, bool FAST, bool WARP = false>
__device__ __forceinline__ float sample_density(const FrameView& f, const MarchParams& mp, f3 p, SampleState<ANISO>& st, LaneCounters& lc,
												uint32_t* __restrict__ list, bool on = true)
{
	if constexpr (ANISO)
	{
#ifdef FM_NO_FMAD
		st.have_grad = false;
		return on ? aniso_density(f, mp, p, st.as, lc) : 0.0f;
#else
		return 0.0f;
#endif
	}
	else
	{
		st.have_grad = WITH_GRAD;
		return eval_density<true, WITH_GRAD, FAST, WARP>(f, p, st.grad, lc, list, on);
	}
}

template <bool ANISO, bool FAST>
__device__ __forceinline__ f3 sample_gradient(const FrameView& f, f3 p, SampleState<ANISO>& st, LaneCounters& lc,
											  uint32_t* __restrict__ list)
{
	if constexpr (ANISO)
	{
#ifdef FM_NO_FMAD
		return st.have_grad ? st.grad : aniso_gradient(f, p, st.as);
#else
		return mk3(0.0f, 0.0f, 0.0f);
#endif
	}
	else
	{
		if (!st.have_grad) eval_density<false, true, FAST>(f, p, st.grad, lc, list);
		return st.grad;
	}
}

// sampleFloor (composition.frag:37-46)
__device__ __forceinline__ void sample_floor(f3 a, f3 r, float out[4])
{
	float const sN = subr(-1.0f, a.y);                         // FLOOR_HEIGHT - a.y
	// (quotients by one divisor share its reciprocal, divs3_shared: the same correctly rounded results; a division by
	// 2 or 4 is the multiplication by 0.5 or 0.25, exactly)
	f3 const b = add3(a, divs3_shared(scale3(r, sN), r.y));    // a + r * (..) / r.y, left to right
	float const mx = subr(b.x, mulr(2.0f, floorf(mulr(b.x, 0.5f))));   // mod(b.x, 2)
	float const mz = subr(b.z, mulr(2.0f, floorf(mulr(b.z, 0.5f))));
	float const fx = (1.0f < mx) ? 0.0f : 1.0f;                // step(m, 1)
	float const fy = (1.0f < mz) ? 0.0f : 1.0f;
	float const g = addr(0.25f, mulr(addr(fx, fy), 0.25f));
	out[0] = g; out[1] = g; out[2] = g; out[3] = 0.5f;
}

__device__ __forceinline__ uint32_t unorm8(float x)
{
	if (!(x > 0.0f)) return 0u;      // also NaN
	if (x >= 1.0f) return 255u;
	return (uint32_t)(addr(mulr(x, 255.0f), 0.5f));
}

// linear -> sRGB, applied by the B8G8R8A8_SRGB swapchain attachment on write (RendererInit2.cpp:50)
__device__ __forceinline__ float srgb_encode(float c)
{
	if (!(c > 0.0f)) return 0.0f;
	if (c >= 1.0f) return 1.0f;
	// __powf = ex2(y * lg2(x)) on the SFU: relative error ~1e-6, i.e. < 3e-4 of an 8-bit code value
	return c <= 0.0031308f ? mulr(12.92f, c) : subr(mulr(1.055f, __powf(c, 1.0f / 2.4f)), 0.055f);
}

// composition.frag:70-122 for one pixel
__device__ __forceinline__ uchar4 shade_pixel(const MarchParams& mp, int px, int py, float4 P, float4 N)
{
	float const u = divr(addr((float)px, 0.5f), (float)mp.W);   // fullscreen.vert UV at the pixel centre
	float const v = divr(addr((float)py, 0.5f), (float)mp.H);
	// viewRay() (composition.frag:59-66): a far-plane POINT used as a direction
	float wh[4];
	mat4_mul_vec4(mp.ipv, subr(mulr(2.0f, u), 1.0f), subr(mulr(2.0f, v), 1.0f), 1.0f, 1.0f, wh);
	f3 const view_ray = divs3_shared(mk3(wh[0], wh[1], wh[2]), wh[3]);
	f3 const cam = mk3(mp.cam[0], mp.cam[1], mp.cam[2]);
	float color[4];
	if (P.w == 0.0f)
	{
		float fl[4];
		sample_floor(cam, view_ray, fl);
#pragma unroll
		for (int k = 0; k < 4; k++) color[k] = mulr(0.75f, fl[k]);
	}
	else
	{
		f3 const world = mk3(P.x, P.y, P.z);
		f3 const normal = mk3(N.x, N.y, N.z);
		// refract(I, N, eta): k = 1 - eta^2 (1 - dot(N,I)^2); k < 0 ? 0 : eta*I - (eta*dot(N,I) + sqrt(k))*N
		f3 const I = normalize3(view_ray);
		float const eta = 1.333f;
		float const dotNI = dot3(normal, I);
		float const k = subr(1.0f, mulr(mulr(eta, eta), subr(1.0f, mulr(dotNI, dotNI))));
		f3 refracted = mk3(0.0f, 0.0f, 0.0f);
		if (k >= 0.0f) refracted = sub3(scale3(I, eta), scale3(normal, addr(mulr(eta, dotNI), sqrtr(k))));
		float fl[4];
		sample_floor(world, refracted, fl);
		float const fv = -dot3(mk3(mp.dir[0], mp.dir[1], mp.dir[2]), normal);
		float const amb = subr(addr(0.15f, 1.0f), mulr(fv, fv));
		float const diffuse[4] = { 120.0f / 255.0f, 185.0f / 255.0f, 255.0f / 255.0f, 255.0f / 255.0f };
#pragma unroll
		for (int kk = 0; kk < 4; kk++) color[kk] = addr(mulr(fv, fl[kk]), mulr(amb, diffuse[kk]));
	}
	uint32_t const r8 = unorm8(srgb_encode(color[0]));
	bool const grey = color[1] == color[0] && color[2] == color[0];   // the floor is grey: one transfer-curve evaluation
	uint32_t const g8 = grey ? r8 : unorm8(srgb_encode(color[1]));
	uint32_t const b8 = grey ? r8 : unorm8(srgb_encode(color[2]));
	return make_uchar4((unsigned char)r8, (unsigned char)g8, (unsigned char)b8, (unsigned char)unorm8(color[3]));
}

// The colour of an uncovered pixel is one of three greys -- 0.75 * sampleFloor(camera, viewRay), the checker of
// composition.frag:37-46 -- decided by which square of the floor the view ray meets.  Four out of five pixels are such
// pixels, and shade_pixel spends ~200 instructions on the reference's exact arithmetic to find that square.  Here the
// meeting point is computed approximately (FMAs, no division by w: the point does not depend on the ray's scale) with a
// bound on how far it can be from shade_pixel's: a point farther than that from every edge of its square lies in the
// same square for both, and the pixel takes its grey from the table `codes` (the three values shade_pixel produces,
// computed by it); a point near an edge, or anything not finite, returns false and shade_pixel decides.
__device__ __forceinline__ bool background_fast(const MarchParams& mp, int px, int py, const uint32_t* codes, uchar4& out)
{
	float const nx = fmaf((float)px + 0.5f, mp.two_w_inv, -1.0f), ny = fmaf((float)py + 0.5f, mp.two_h_inv, -1.0f);
	const float* m = mp.ipv;
	// (mp.bg_c[r] = m[8 + r] + m[12 + r]: one rounding, the same one shade_pixel makes -- the cancellation of the z row is
	// common to both)
	float const wx = fmaf(m[0], nx, fmaf(m[4], ny, mp.bg_c[0]));
	float const wy = fmaf(m[1], nx, fmaf(m[5], ny, mp.bg_c[1]));
	float const wz = fmaf(m[2], nx, fmaf(m[6], ny, mp.bg_c[2]));
	// mp.bg_err: how far wx, wy, wz can be from shade_pixel's -- 64 ulps of the largest |m[r]| + |m[4 + r]| + |bg_c[r]| cover
	// the handful of roundings that differ; the quotient and the products add relative errors of a few ulps
	float const ry = __fdividef(1.0f, wy);
	float const t = (-1.0f - mp.cam[1]) * ry;
	float const bx = fmaf(wx, t, mp.cam[0]), bz = fmaf(wz, t, mp.cam[2]);
	float const e = mp.bg_err * fabsf(t) * (1.0f + (fabsf(wx) + fabsf(wz)) * fabsf(ry)) + 1e-5f * (fabsf(bx) + fabsf(bz) + mp.bg_cam);
	float const mx = bx - 2.0f * floorf(bx * 0.5f), mz = bz - 2.0f * floorf(bz * 0.5f);       // mod(b, 2) in [0, 2)
	float const d = fminf(fminf(fminf(mx, fabsf(mx - 1.0f)), 2.0f - mx), fminf(fminf(mz, fabsf(mz - 1.0f)), 2.0f - mz));
	if (!(d > e && e < 0.25f)) return false;
	uint32_t const k = (mx < 1.0f ? 1u : 0u) + (mz < 1.0f ? 1u : 0u);                        // step(m, 1) summed
	uint32_t const c = codes[k];
	out = make_uchar4((unsigned char)c, (unsigned char)c, (unsigned char)c, (unsigned char)codes[3]);
	return true;
}

__device__ __forceinline__ bool pixel_active(const MarchParams& mp, int px, int py)
{
	bool active = px < mp.W && py < mp.H;
	if (active && mp.rx1 > 0) active = px >= mp.rx0 && px < mp.rx1 && py >= mp.ry0 && py < mp.ry1;
	if (active && mp.part_world > 1)
	{
		int const tile = (py / mp.part_th) * mp.part_tiles_x + (px / mp.part_tw);
		active = (tile % mp.part_world) == mp.part_rank;
	}
	uint32_t const index = (uint32_t)py * (uint32_t)mp.W + (uint32_t)px;
	if (active && mp.skip_last_pixel && index == (uint32_t)mp.W * (uint32_t)mp.H - 1u) active = false;   // ThreadPool.cpp:50
	return active;
}

// every pixel: background + work list of the 8x4 tiles that hold covered pixels.  CTA = 4x2 tiles = 32x8 pixels.
// what: 1 = the tile list, 2 = the outputs of the uncovered pixels, 3 = both.  With stage timing off the second half runs
// on the context's side stream beside k_march_long (launch_march).
__global__ void __launch_bounds__(256) k_classify(MarchParams mp, const float* __restrict__ depth,
												  float4* __restrict__ pos_out, float4* __restrict__ nrm_out,
												  uchar4* __restrict__ rgba_out, uint32_t* __restrict__ tiles,
												  uint32_t* __restrict__ n_tiles, int what)
{
	pdl_enter();
	int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	// the three greys of the floor + its alpha as shade_pixel encodes them (background_fast)
	__shared__ uint32_t s_bg[4];
	if ((what & 2) && mp.do_march && mp.do_shade)
	{
		if (threadIdx.x < 4)
		{
			float const g = 0.25f + (float)threadIdx.x * 0.25f;        // 0.25 + (fx + fy) / 4, exact
			s_bg[threadIdx.x] = threadIdx.x < 3 ? unorm8(srgb_encode(mulr(0.75f, g))) : unorm8(mulr(0.75f, 0.5f));
		}
		__syncthreads();
	}
	// (under a region partition the grid covers the region only: its bounds are multiples of 64 pixels)
	int const tx = (mp.rx0 >> 3) + blockIdx.x * 4 + (warp & 3), ty = (mp.ry0 >> 2) + blockIdx.y * 2 + (warp >> 2);
	int const px = tx * 8 + (lane & 7), py = ty * 4 + (lane >> 3);
	bool const active = pixel_active(mp, px, py);
	uint32_t const index = (uint32_t)py * (uint32_t)mp.W + (uint32_t)px;
	bool covered = false;
	if (active)
	{
		if (mp.do_march)
		{
			covered = depth[index] != 1.0f;     // `if (z == 1.0f) return;` (RayMarcher.cpp:264)
			if (!covered && (what & 2))
			{
				float4 const zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);   // RayMarcher.cpp:262-263
				pos_out[index] = zero;
				nrm_out[index] = zero;
				if (mp.do_shade)
				{
					uchar4 c;
					if (!mp.bg_fast || !background_fast(mp, px, py, s_bg, c)) c = shade_pixel(mp, px, py, zero, zero);
					rgba_out[index] = c;
				}
			}
		}
		else rgba_out[index] = shade_pixel(mp, px, py, pos_out[index], nrm_out[index]);   // shade-only pass
	}
	if (!mp.do_march || !(what & 1)) return;
	__shared__ uint32_t s_any[8];
	__shared__ uint32_t s_base;
	bool const any = __any_sync(0xffffffffu, covered);
	if (lane == 0) s_any[warp] = any ? 1u : 0u;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		uint32_t c = 0;
#pragma unroll
		for (int w = 0; w < 8; w++) c += s_any[w];
		s_base = c ? atomicAdd(n_tiles, c) : 0u;
	}
	__syncthreads();
	if (any && lane == 0)
	{
		uint32_t slot = s_base;
		for (int w = 0; w < warp; w++) slot += s_any[w];
		tiles[slot] = ((uint32_t)ty << 16) | (uint32_t)tx;
	}
}

__device__ __forceinline__ void add_counts(LaneCounters& a, const LaneCounters& b)
{
	a.candidates += b.candidates; a.neighbours += b.neighbours; a.overflow += b.overflow;
}

// Per ray: what the empty-space skip needs of the (constant) step.  rcp = the refined reciprocal div2_shared derives
// from a divisor (same three operations, so the quotients below are div2_shared's bit for bit); ok = every component
// is a normal number in div2_shared's range (no zero, NaN, infinity).
#ifdef FM_LONG_PROFILE
#define FM_PROF_GENERAL() (skips += 0x10000u)      // profiling build: general-path skips ride in the high half of `skips`
#else
#define FM_PROF_GENERAL() ((void)0)
#endif
struct StepInfo
{
	f3 rcp;
	bool ok;              // every component is +-0 or a normal number in div2_shared's range
	bool px, py, pz;      // the component is > 0: the far plane of a cell on that axis is its Max
	bool zx, zy, zz;      // the component is exactly +-0 (the ray through the image centre line of a symmetric camera)
	bool any_zero;
};
__device__ __forceinline__ float refined_rcp(float s)
{
	float r0;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(s));
	return fmaf(r0, fmaf(-s, r0, 1.0f), r0);
}
__device__ __forceinline__ StepInfo step_info(f3 step)
{
	StepInfo si;
	float const ax = fabsf(step.x), ay = fabsf(step.y), az = fabsf(step.z);
	si.px = step.x > 0.0f; si.py = step.y > 0.0f; si.pz = step.z > 0.0f;
	si.zx = step.x == 0.0f; si.zy = step.y == 0.0f; si.zz = step.z == 0.0f;
	si.any_zero = si.zx | si.zy | si.zz;
	si.ok = (si.zx | ((ax >= 0x1p-60f) & (ax <= 0x1p60f))) & (si.zy | ((ay >= 0x1p-60f) & (ay <= 0x1p60f))) &
		(si.zz | ((az >= 0x1p-60f) & (az <= 0x1p60f)));
	si.rcp = mk3(refined_rcp(step.x), refined_rcp(step.y), refined_rcp(step.z));
	return si;
}

// the general intersectAABB, kept out of line: the walk below takes it only for degenerate steps and samples that
// rounding put outside their own cell
__device__ __noinline__ f3 skip_cell_general(f3 o, f3 d, f3 nmin, float cw)
{
	f3 const nmax = add3(nmin, mk3(cw, cw, cw));
	return add3(intersect_aabb(o, d, nmin, nmax), d);
}

// one `position += step` with the empty-space skip of RayMarcher.cpp:279-306 (skips do not consume MaxSteps).
// Returns true when the ray has left the density grid for good, i.e. this and all later samples are misses:
// outside the grid every particle is farther than h (the grid is the particle AABB padded by h), so the
// density is 0 there; each coordinate moves monotonically, so a ray that is outside on an axis and moving
// away on it can never come back.  Stopping there leaves the result unchanged.
//
// The skip is a serial chain -- each exit point is computed from the previous one in FP32 -- and the longest such
// chain is what k_march_long takes.  intersectAABB (RayMarcher.cpp:51-62) forms six quotients, but for a sample inside
// its cell only three can win: with d > 0 on an axis (boxMax - o)/d >= (boxMin - o)/d (RN subtraction and division
// are monotone), so t2 on that axis is the quotient of the FAR plane, and t1 <= 0 on every axis whose near plane
// is not ahead of the sample.  Then tNear <= 0, and whenever tFar > 0, max(tNear, tFar) = tFar = the smallest of
// the three far quotients, value and bits (equal positive values have equal bits).  An axis with d = +-0 (the centre
// column / row of a symmetric camera: r02q, 88 general-path skips at ~1500 cycles each were C3's slowest ray) has
// quotients -inf / +inf strictly between its planes and constrains nothing.  Anything else -- a subnormal or
// non-finite step component, a sample rounding put outside its box or exactly on a plane of a d = 0 axis, tFar <= 0,
// a numerator outside div2_shared's range -- takes the general function.
//
// occ_s != 0: the occupancy bitmap has been staged in shared memory at that (shared-window) address.  Every iteration
// of the chain needs its cell's bit before it can go on; out of L2 that load IS the chain (ncu / clock64 r02o: ~900
// cycles per skipped cell at C3, 72 of the kernel's 85 us in the walk of its slowest ray).
// The skip of one cell without a branch: writes the exit point + step into `next` and returns whether the shortcut
// applies (else the caller takes the general function).  ZERO = the step may have +-0 components.
template <bool ZERO>
__device__ __forceinline__ bool skip_cell_fast(const StepInfo& si, f3 step, f3 nmin, f3 nmax, f3 p, f3& next)
{
	float const nfx = subr(si.px ? nmax.x : nmin.x, p.x);
	float const nfy = subr(si.py ? nmax.y : nmin.y, p.y);
	float const nfz = subr(si.pz ? nmax.z : nmin.z, p.z);
	// the near plane is not ahead of the sample:  d > 0: Min <= p;  d < 0: p <= Max
	bool ok = ((si.px ? nmin.x : p.x) <= (si.px ? p.x : nmax.x)) & ((si.py ? nmin.y : p.y) <= (si.py ? p.y : nmax.y)) &
		((si.pz ? nmin.z : p.z) <= (si.pz ? p.z : nmax.z));
	float q0 = mulr(nfx, si.rcp.x);
	float qx = fmaf(si.rcp.x, fmaf(-step.x, q0, nfx), q0);
	q0 = mulr(nfy, si.rcp.y);
	float qy = fmaf(si.rcp.y, fmaf(-step.y, q0, nfy), q0);
	q0 = mulr(nfz, si.rcp.z);
	float qz = fmaf(si.rcp.z, fmaf(-step.z, q0, nfz), q0);
	float mx = fabsf(nfx), my = fabsf(nfy), mz = fabsf(nfz);
	float const inf = __int_as_float(0x7f800000);
	if (ZERO)
	{
		// an axis the ray does not move on (d = +-0): strictly between the planes both quotients are -inf / +inf,
		// t1 = -inf, t2 = +inf, the axis constrains nothing; ON a plane 0/0 = NaN enters glm's min / max: general path
		ok = ok & si.ok & (!si.zx | ((nmin.x < p.x) & (p.x < nmax.x))) & (!si.zy | ((nmin.y < p.y) & (p.y < nmax.y))) &
			(!si.zz | ((nmin.z < p.z) & (p.z < nmax.z)));
		qx = si.zx ? inf : qx; qy = si.zy ? inf : qy; qz = si.zz ? inf : qz;
		mx = si.zx ? 1.0f : mx; my = si.zy ? 1.0f : my; mz = si.zz ? 1.0f : mz;
	}
	float const lo = fminf(fminf(mx, my), mz), hi = fmaxf(fmaxf(mx, my), mz);
	float const t_far = fminf(fminf(qx, qy), qz);
	ok = ok & (lo >= 0x1p-60f) & (hi <= 0x1p60f) & (t_far > 0.0f);
	if (ZERO) ok = ok & (t_far < inf);
	next = add3(add3(p, scale3(step, t_far)), step);
	return ok;
}

// the test at the head of advance's loop for a sample at p: inside the density grid, in a cell whose flag is set
// (QueryDensityGrid, Dataset.cpp:26-47; `node->Flag`, RayMarcher.cpp:294) -- advance returns such a sample as it is
__device__ __forceinline__ bool in_occupied_cell(const FrameView& f, f3 p, uint32_t occ_s)
{
	float const fx = floorf(mulr(subr(p.x, f.mn.x), f.inv_cell_width.x));
	float const fy = floorf(mulr(subr(p.y, f.mn.y), f.inv_cell_width.y));
	float const fz = floorf(mulr(subr(p.z, f.mn.z), f.inv_cell_width.z));
	bool const inside = (fx >= 0.0f) & (fy >= 0.0f) & (fz >= 0.0f) & (fx < (float)f.gdim.x) & (fy < (float)f.gdim.y) & (fz < (float)f.gdim.z);
	if (!inside) return false;
	uint32_t const c = (uint32_t)(int)fx + (uint32_t)f.gdim.x * ((uint32_t)(int)fy + (uint32_t)f.gdim.y * (uint32_t)(int)fz);
	uint32_t const word = occ_s ? lds_u32(occ_s + ((c >> 5) << 2)) : __ldg(f.occ_bits + (c >> 5));
	return ((word >> (c & 31u)) & 1u) != 0u;
}

template <bool ZERO>
__device__ __forceinline__ bool advance_t(const FrameView& f, const MarchParams& mp, f3 step, const StepInfo& si, f3& position,
										  f3& prev, uint32_t& skips, uint32_t occ_s)
{
	prev = position;
	position = add3(position, step);
	bool inside;
	for (;;)
	{
		// Frame::QueryDensityGrid (Dataset.cpp:26-47), as density_cell_of
		float const fx = floorf(mulr(subr(position.x, f.mn.x), f.inv_cell_width.x));
		float const fy = floorf(mulr(subr(position.y, f.mn.y), f.inv_cell_width.y));
		float const fz = floorf(mulr(subr(position.z, f.mn.z), f.inv_cell_width.z));
		// (NaN -> outside, as the reference's int conversion)
		inside = (fx >= 0.0f) & (fy >= 0.0f) & (fz >= 0.0f) & (fx < (float)f.gdim.x) & (fy < (float)f.gdim.y) & (fz < (float)f.gdim.z);
		if (!inside) break;
		uint32_t const c = (uint32_t)(int)fx + (uint32_t)f.gdim.x * ((uint32_t)(int)fy + (uint32_t)f.gdim.y * (uint32_t)(int)fz);
		uint32_t const word = occ_s ? lds_u32(occ_s + ((c >> 5) << 2)) : __ldg(f.occ_bits + (c >> 5));
		// node->Min = m_Min + vec3(x,y,z)*cellWidth; node->Max = Min + vec3(cellWidth) (Dataset.cpp:132-133); fx is the
		// integer the reference converts back to float.  (The exit point is computed while the bitmap word is on its way.)
		f3 const nmin = add3(mk3(f.mn.x, f.mn.y, f.mn.z), scale3(mk3(fx, fy, fz), f.cell_width));
		f3 const nmax = add3(nmin, mk3(f.cell_width, f.cell_width, f.cell_width));
		f3 next;
		bool const fast = skip_cell_fast<ZERO>(si, step, nmin, nmax, position, next);
		if ((word >> (c & 31u)) & 1u) break;
		prev = position;
		if (fast) position = next;
		else { position = skip_cell_general(position, step, nmin, f.cell_width); FM_PROF_GENERAL(); }
		skips++;
	}
	if (!inside && mp.early_out)
	{
		float const rx = floorf(mulr(subr(position.x, f.mn.x), f.inv_cell_width.x));
		float const ry = floorf(mulr(subr(position.y, f.mn.y), f.inv_cell_width.y));
		float const rz = floorf(mulr(subr(position.z, f.mn.z), f.inv_cell_width.z));
		float const m = mp.early_margin;
		return (rx < -m && step.x <= 0.0f) || (rx >= (float)f.gdim.x + m && step.x >= 0.0f) ||
			   (ry < -m && step.y <= 0.0f) || (ry >= (float)f.gdim.y + m && step.y >= 0.0f) ||
			   (rz < -m && step.z <= 0.0f) || (rz >= (float)f.gdim.z + m && step.z >= 0.0f);
	}
	return false;
}

__device__ __forceinline__ bool advance(const FrameView& f, const MarchParams& mp, f3 step, const StepInfo& si, f3& position,
										f3& prev, uint32_t& skips, uint32_t occ_s = 0u)
{
	return (si.ok & !si.any_zero) ? advance_t<false>(f, mp, step, si, position, prev, skips, occ_s)
								  : advance_t<true>(f, mp, step, si, position, prev, skips, occ_s);
}

// the sample at `position` reached the threshold (RayMarcher.cpp:327-341 / :405-417): optional bisection, normal
template <bool ANISO, bool FAST>
__device__ __forceinline__ void finish_hit(const FrameView& f, const MarchParams& mp, f3 prev, f3 position,
										   SampleState<ANISO>& st, LaneCounters& lc, uint32_t* __restrict__ list, float4& P, float4& N)
{
	// optional refinement (north_star item 3; not in the reference): bisect between the position before the last
	// advance and the hit sample.  That position was never sampled when the hit is the ray's first sample or follows
	// an empty-cell jump, so it is sampled first: without a sign change there is no bracket and the hit stays where
	// the reference puts it.
	f3 lo = prev, hi = position;
	if (mp.bisection_steps > 0)
	{
		SampleState<ANISO> tmp;
		float const d_lo = sample_density<ANISO, false, false>(f, mp, lo, tmp, lc, list);
		lc.steps++;
		if (d_lo < mp.iso)
			for (int b = 0; b < mp.bisection_steps; b++)
			{
				f3 const mid = scale3(add3(lo, hi), 0.5f);
				float const dm = sample_density<ANISO, false, false>(f, mp, mid, tmp, lc, list);
				lc.steps++;
				if (dm >= mp.iso) { hi = mid; st = tmp; } else lo = mid;
			}
	}
	P = make_float4(hi.x, hi.y, hi.z, 1.0f);
	f3 const n = normalize3(sample_gradient<ANISO, FAST>(f, hi, st, lc, list));   // glm::normalize(normal) (RayMarcher.cpp:338)
	N = make_float4(n.x, n.y, n.z, 1.0f);
	lc.hits++;
}

// ---- ray queues between the three march phases -----------------------------------------------------------
// A ray that is not finished by a phase is handed on as 32 bytes: (position.xyz, pixel index) (step.xyz, samples taken)
__device__ __forceinline__ void push_rays(bool want, float4* __restrict__ q, uint32_t* __restrict__ n, uint32_t index,
										  f3 position, f3 step, int i)
{
	uint32_t const m = __ballot_sync(0xffffffffu, want);
	if (m == 0u) return;
	int const lane = threadIdx.x & 31;
	uint32_t base = 0;
	if (lane == __ffs(m) - 1) base = atomicAdd(n, (uint32_t)__popc(m));
	base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
	if (want)
	{
		uint32_t const slot = base + __popc(m & ((1u << lane) - 1u));
		q[2 * (size_t)slot] = make_float4(position.x, position.y, position.z, __uint_as_float(index));
		q[2 * (size_t)slot + 1] = make_float4(step.x, step.y, step.z, __int_as_float(i));
	}
}

__device__ __forceinline__ void write_pixel(const MarchParams& mp, uint32_t index, float4 P, float4 N,
											float4* __restrict__ pos_out, float4* __restrict__ nrm_out,
											uchar4* __restrict__ rgba_out)
{
	pos_out[index] = P;
	nrm_out[index] = N;
	if (mp.do_shade) rgba_out[index] = shade_pixel(mp, (int)(index % (uint32_t)mp.W), (int)(index / (uint32_t)mp.W), P, N);
}

template <bool FIRST = false>
__device__ __forceinline__ void flush_counters(const LaneCounters& lc, DeviceCounters* __restrict__ counters)
{
	// per-warp counter reduction, one atomic per counter per warp; k_march_first also books its candidates as
	// DeviceCounters::first_candidates (slot 8), the candidates its staged walk examined (9) and its fallbacks (10)
	constexpr int N = FIRST ? 11 : 8;
	uint32_t vals[11] = { lc.covered, lc.hits, lc.steps, lc.skips, lc.candidates, lc.neighbours, lc.early_exits, lc.overflow, lc.candidates,
						  lc.examined, lc.fallbacks };
	unsigned long long* dst = reinterpret_cast<unsigned long long*>(counters);
#pragma unroll
	for (int k = 0; k < N; k++)
	{
		uint32_t const s = __reduce_add_sync(0xffffffffu, vals[k]);
		if ((threadIdx.x & 31) == 0 && s) atomicAdd(dst + k, (unsigned long long)s);
	}
}

// phase A: warp = one covered 8x4 tile, lanes = rays, the FIRST sample of every ray with density and gradient sums
// together (the depth pre-pass seeds the ray right in front of the surface, so ~95% of the rays end here)
// isotropic k_march_first: FM_FIRST_STAGED = 1: FM_FIRST_WARPS warps per CTA, one WarpStage of dynamic shared memory per warp;
// 0: 8 warps per CTA, dynamic shared memory = the in-range lists of the global-memory walk.  Anisotropic: 8 warps, none.
#ifndef FM_FIRST_STAGED
#define FM_FIRST_STAGED 0
#endif
constexpr int kFirstThreads = FM_FIRST_STAGED ? FM_FIRST_WARPS * 32 : 256;
constexpr size_t kFirstWarpBytes = FM_FIRST_STAGED ? sizeof(WarpStage) : (size_t)kListWords * 4;      // dynamic shared memory per warp
constexpr size_t kFirstSmem = (size_t)(kFirstThreads / 32) * kFirstWarpBytes;
constexpr size_t kLongSmem = (size_t)8 * kListWords * 4;           // k_march_long, isotropic: the lists of its 8 warps
static_assert(kStageCap * 16 <= 65536, "stage byte offsets are kept in 16 bits");
static_assert(sizeof(WarpStage) >= kListWords * 4, "the global-memory walk's list lives in the stage");

// (The view is a kernel parameter: its fields reach the instructions as uniform constant-bank operands.  Reading it
// from device memory instead -- which would let the march be queued before the host knows the frame's grid parameters
// -- was measured, r02c-e: through a __constant__ table indexed per context k_march_first +20 %, through a shared-memory
// copy +12 %, k_march_long +9 %.  The host gets the parameters early on a side stream instead, fm_grid.cu.)
#ifdef FM_FIRST_PROFILE
// profiling build (tools/build_variant.sh x -DFM_FIRST_PROFILE, read by tools/first_profile.py through fr_debug_first_profile):
// per warp: start, finish (globaltimer ns), duration of its longest tile, (max skips of a lane on that tile << 32 | tiles done)
__device__ unsigned long long g_first_prof[4 * 16384];
__device__ __forceinline__ unsigned long long prof_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
template <bool FAST, bool ANISO>
__global__ void __launch_bounds__(ANISO ? 256 : kFirstThreads, ANISO ? FM_ANISO_MINBLOCKS : (FM_FIRST_STAGED ? FM_FIRST_MINBLOCKS : FM_MARCH_MINBLOCKS)) k_march_first(FrameView f, MarchParams mp, const float* __restrict__ depth,
														 float4* __restrict__ pos_out, float4* __restrict__ nrm_out,
														 uchar4* __restrict__ rgba_out, const uint32_t* __restrict__ tiles,
														 RayQueues rq, DeviceCounters* __restrict__ counters, uint32_t occ_words)
{
	pdl_enter();
	constexpr uint32_t FULL = 0xffffffffu;
	int const lane = threadIdx.x & 31;
	extern __shared__ __align__(16) unsigned char s_dyn[];
	// the global-memory walk (every tile when FM_FIRST_STAGED = 0, else tiles that do not fit the stage and bisection
	// samples) keeps its list at the start of the warp's share
	uint32_t* const list = reinterpret_cast<uint32_t*>(s_dyn) + (ANISO ? 0 : (threadIdx.x >> 5) * (kFirstWarpBytes / 4) + lane);
	LaneCounters lc = {};
	uint32_t const count = __ldcg(rq.ctl + 0);
	// the occupancy bitmap, staged when it fits without costing a resident CTA (see advance, k_march_long)
	uint32_t occ_s = 0u;
	__shared__ __align__(8) unsigned long long s_bitmap_bar;
	bool bitmap_pending = false;
	if (occ_words != 0u && count != 0u)
	{
		uint32_t* const occ = reinterpret_cast<uint32_t*>(s_dyn + (ANISO ? kAnisoFirstSmem : kFirstSmem));
		bitmap_stage_begin(occ, f.occ_bits, occ_words, &s_bitmap_bar);
		bitmap_pending = true;
		occ_s = smem_addr(occ);
	}
	f3 const cam = mk3(mp.cam[0], mp.cam[1], mp.cam[2]);

	// the first tile of a warp is the one with its own number: thousands of warps drawing their first ticket from one
	// counter at the same moment queue up at that address for longer than a tile takes to launch; afterwards the
	// tickets are spread in time
	// (r02w-y: this kernel marching the rays it queues between its tiles -- edge tiles first, a semaphore-published queue,
	// k_march_long for what is left; the code is in the history at d55db80 -- loses: a queued ray's serial walk runs ~4x
	// slower among the 24 busy warps of an SM than among k_march_long's 16 mostly idle ones, C2 0.124 + 0.052 -> 0.33 +
	// 0.04 ms.)
	// (r02z: tiles cost about the same -- C2: p50 27.8 us, p90 30.1 us -- and a warp gets only 3.67 of them, so the kernel
	// lasts 4 tiles while the SMs drain for the last quarter of it.  Letting only ceil(tiles / 4) warps take part, so that
	// every round of tiles is full, does not help: a tile takes as long among 22 warps as among 24 -- 0.1249 -> 0.1268 ms.)
	uint32_t const nwarps = (gridDim.x * blockDim.x) >> 5;
	uint32_t const my_first = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	bool first = true;
#ifdef FM_FIRST_PROFILE
	unsigned long long const prof_start = prof_now();
	unsigned long long prof_longest = 0, prof_info = 0;
	uint32_t prof_tiles = 0;
#endif
	for (;;)
	{
#ifdef FM_FIRST_PROFILE
		unsigned long long const prof_t0 = prof_now();
		uint32_t const prof_s0 = lc.skips;
#endif
		uint32_t t = my_first;
		if (!first)
		{
			if (lane == 0) t = nwarps + atomicAdd(rq.ctl + 1, 1u);
			t = __shfl_sync(FULL, t, 0);
		}
		first = false;
		if (t >= count) break;
		uint32_t const txy = __ldg(tiles + t);
		int const px = (int)(txy & 0xffffu) * 8 + (lane & 7), py = (int)(txy >> 16) * 4 + (lane >> 3);
		uint32_t const index = (uint32_t)py * (uint32_t)mp.W + (uint32_t)px;
		bool covered = pixel_active(mp, px, py);
		float z = 1.0f;
		if (covered) { z = depth[index]; covered = z != 1.0f; }   // uncovered pixels were finished by k_classify

		float4 P = make_float4(0.0f, 0.0f, 0.0f, 0.0f), N = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		f3 position = mk3(0.0f, 0.0f, 0.0f), step = position, prev = position;
		bool more = false, sample = false;
		if (bitmap_pending) { bitmap_stage_wait(&s_bitmap_bar); bitmap_pending = false; }
		if (covered)
		{
			lc.covered++;
			// pixel CORNER, not centre (RayMarcher.cpp:268-270)
			float const cx = subr(mulr((float)px, mp.two_w_inv), 1.0f);
			float const cy = subr(mulr((float)py, mp.two_h_inv), 1.0f);
			float wh[4];
			mat4_mul_vec4(mp.ipv, cx, cy, z, 1.0f, wh);
			position = divs3_shared(mk3(wh[0], wh[1], wh[2]), wh[3]);
			step = scale3(normalize3(sub3(position, cam)), mp.step_size);
			prev = position;
			if (mp.max_steps > 0)
			{
				if (advance(f, mp, step, step_info(step), position, prev, lc.skips, occ_s)) lc.early_exits++;
				else sample = true;
			}
		}
		// The empty-space skip loop of `advance` ends after a different number of iterations per lane, and the compiler
		// only reconverges at the end of the enclosing branch: without this barrier the warp walks the 27 cells below in
		// ~1.8 separate groups of lanes (ncu r01 s6: 16.5 active threads per instruction).  From here on the lanes run in
		// lock step again; lanes without a sample walk empty ranges.
		__syncwarp();
		// isotropic: the gradient sum rides along with the density on this sample (unless bisection moves the hit)
		SampleState<ANISO> st;
		float density = 0.0f;
		bool staged = false;
		if constexpr (!ANISO && FM_FIRST_STAGED)
		{
			if (mp.bisection_steps == 0)
			{
				WarpStage& ws = *reinterpret_cast<WarpStage*>(s_dyn + (threadIdx.x >> 5) * kFirstWarpBytes);
				staged = eval_density_staged<true, FAST>(f, position, sample, ws, density, st.grad, lc);
				st.have_grad = true;
				if (!staged && sample) lc.fallbacks++;
				__syncwarp();
			}
		}
		if constexpr (ANISO && FM_ANISO_FIRST_LIST)
		{
#ifdef FM_NO_FMAD
			// the tile's first samples share one candidate list (aniso_list_build); the hits' gradient pass reuses it
			uint32_t* const wl = reinterpret_cast<uint32_t*>(s_dyn) + (threadIdx.x >> 5) * kAnisoListCap;
			uint32_t wn;
			bool listed;
			st.have_grad = false;
			density = aniso_density_warp(f, mp, position, sample, st.as, lc, wl, wn, listed);
			bool const hit = sample && density >= mp.iso && mp.bisection_steps == 0;
			if (listed && __any_sync(0xffffffffu, hit))
			{
				st.grad = aniso_gradient_list(f, position, hit, st.as, wl, wn);
				st.have_grad = hit;
			}
#endif
		}
		else if (!staged)
			density = (!ANISO && mp.bisection_steps == 0)
				? sample_density<ANISO, true, FAST, true>(f, mp, position, st, lc, list, sample)
				: sample_density<ANISO, false, false, true>(f, mp, position, st, lc, list, sample);
		__syncwarp();
		if (sample)
		{
			lc.steps++;
			if (density >= mp.iso) finish_hit<ANISO, FAST>(f, mp, prev, position, st, lc, list, P, N);   // RayMarcher.cpp:327
			else if (mp.max_steps > 1)
			{
				// most rays that miss here are silhouette rays about to leave the grid: settle them now
				f3 p2 = position, prev2 = position;
				uint32_t skips2 = 0;
				if (advance(f, mp, step, step_info(step), p2, prev2, skips2, occ_s)) { lc.skips += skips2; lc.early_exits++; }
				else more = true;      // (the queue keeps the state before this advance)
			}
		}
		__syncwarp();
		push_rays(more, rq.q1, rq.ctl + 2, index, position, step, 1);
		if (covered && !more) write_pixel(mp, index, P, N, pos_out, nrm_out, rgba_out);
#ifdef FM_FIRST_PROFILE
		{
			__syncwarp();
			unsigned long long const dt = prof_now() - prof_t0;
			uint32_t sk = lc.skips - prof_s0;
			for (int o = 16; o; o >>= 1) sk = max(sk, __shfl_xor_sync(FULL, sk, o));
			prof_tiles++;
			if (dt > prof_longest) { prof_longest = dt; prof_info = ((unsigned long long)sk << 32) | ((unsigned long long)(txy >> 16) << 16) | (txy & 0xffffu); }
		}
#endif
	}
	// (r03s: the warps that are out of tiles writing the outputs of the uncovered pixels -- to fill the drain -- is a loss:
	// 24 warps per SM at 72 registers write 60 MB far slower than k_classify's 64: k_march_first 0.123 -> 0.158 ms.)
#ifdef FM_FIRST_PROFILE
	if (lane == 0)
	{
		uint32_t const w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
		if (w < 16384u)
		{
			g_first_prof[4 * w + 0] = prof_start; g_first_prof[4 * w + 1] = prof_now();
			g_first_prof[4 * w + 2] = (prof_longest << 16) | prof_tiles; g_first_prof[4 * w + 3] = prof_info;
		}
	}
#endif
	flush_counters<true>(lc, counters);
}

// ---- one sample evaluated by the whole warp: LANES = CANDIDATES ----------------------------------------------------------
// The 9 ranges of the query are one flat index space, 32 candidates per load instruction; the in-range ones are
// compacted in order (ballot), their W / gradW are computed one per lane, and only the sums themselves run as the
// reference's ordered chain: the same operations in the same order as eval_density, hence the same bits, at ~1/10 of
// the latency of one lane doing it alone.  k_march_long uses it for the gradient sum of a hit (one lane of the warp
// used to re-walk the 27 cells alone while 31 waited: 6 % of the kernel's stall samples, at the end of every hitting
// ray's critical path).  (A whole kernel built on it -- one queued ray per CTA, 8 warps x 4 samples of a window each --
// was measured, r02f: per-ray latency fell by more than half, but with ~1 700 rays and one CTA per ray the rays no
// longer run all at once: C2 0.062 -> 0.105 ms, C3 0.080 -> 0.222 ms.  More than half of k_march_long is the serial
// walk through empty cells (ncu: 56 % of its instructions), which no amount of evaluation parallelism shortens.)
#ifndef FM_COOP_CAP
#define FM_COOP_CAP 128                   // in-range candidates a warp collects before it sums them
#endif
constexpr int kCoopCap = FM_COOP_CAP;

struct CoopWarp
{
	float4 ent[kCoopCap];                 // (d0, d1, d2, l2) of the in-range candidates, in the reference's list order
	float4 con[kCoopCap];                 // their contributions: (W, gradW) or, fast normals, (W, coefficient)
};

static_assert(sizeof(CoopWarp) <= (size_t)kListWords * 4, "k_march_long evaluates the hit gradient in its warp's list area");

template <bool FAST>
__device__ __forceinline__ void coop_flush(const FrameView& f, CoopWarp& cw, uint32_t cnt, uint32_t& nn, float& density, f3& g)
{
	int const lane = threadIdx.x & 31;
	__syncwarp();
	for (uint32_t k = lane; k < cnt; k += 32)
	{
		float4 const e = cw.ent[k];
		if (!FAST)
		{
			float W;
			f3 gw;
			spline_W_gradW_inrange(f.kernel, mk3(-e.x, -e.y, -e.z), e.w, W, gw);
			cw.con[k] = make_float4(W, gw.x, gw.y, gw.z);
		}
		else cw.con[k] = make_float4(spline_W_inrange(f.kernel, e.w), spline_gradW_coeff_fast(f.kernel, e.w, mulr(sqrtr(e.w), f.kernel.h_inv)), 0.0f, 0.0f);
	}
	__syncwarp();
	// the sums, in list order (every lane runs the chain: the result is warp-uniform without a broadcast)
#pragma unroll 1
	for (uint32_t k = 0; k < cnt; k++)
	{
		if (nn < (uint32_t)kMaxNeighbors)   // list truncation of RayMarcher.cpp:312
		{
			float4 const c = cw.con[k];
			if (!FAST) g = add3(g, mk3(c.y, c.z, c.w));
			else
			{
				float4 const e = cw.ent[k];
				g.x = fmaf(c.y, -e.x, g.x); g.y = fmaf(c.y, -e.y, g.y); g.z = fmaf(c.y, -e.z, g.z);
			}
			density = addr(density, c.x);
		}
		nn++;
	}
	__syncwarp();
}

// density and gradient sum at p, the whole warp on one sample; warp-uniform results
template <bool FAST>
__device__ __forceinline__ void coop_eval(const FrameView& f, f3 p, CoopWarp& cw, float& density_out, f3& grad_out, uint32_t& cand_out,
										  uint32_t& nn_out)
{
	constexpr uint32_t FULL = 0xffffffffu;
	int const lane = threadIdx.x & 31;
	int const kx = search_cell_of(f.search_inv, p.x) - f.kmin.x;
	int const ky = search_cell_of(f.search_inv, p.y) - f.kmin.y;
	int const kz = search_cell_of(f.search_inv, p.z) - f.kmin.z;
	int const z0 = max(kz - 1, 0), z1 = min(kz + 1, f.kdim.z - 1);
	// lane r < 9: range r of the query (x, then y; three z cells = one contiguous particle range)
	uint32_t rb = 0, rn = 0;
	if (lane < 9 && z0 <= z1)
	{
		int const x = kx + lane / 3 - 1, y = ky + lane % 3 - 1;
		if ((unsigned)x < (unsigned)f.kdim.x && (unsigned)y < (unsigned)f.kdim.y)
		{
			uint32_t const base = ((uint32_t)x * (uint32_t)f.kdim.y + (uint32_t)y) * (uint32_t)f.kdim.z;
			rb = __ldg(f.cell_start + base + z0);
			rn = __ldg(f.cell_start + base + z1 + 1) - rb;
		}
	}
	uint32_t inc = rn;
#pragma unroll
	for (int o = 1; o < 16; o <<= 1)
	{
		uint32_t const t = __shfl_up_sync(FULL, inc, o);
		if (lane >= o) inc += t;
	}
	uint32_t const total = __shfl_sync(FULL, inc, 8);
	uint32_t ro[9], rnn[9], rbb[9];
#pragma unroll
	for (int r = 0; r < 9; r++)
	{
		rnn[r] = __shfl_sync(FULL, rn, r);
		ro[r] = __shfl_sync(FULL, inc, r) - rnn[r];
		rbb[r] = __shfl_sync(FULL, rb, r);
	}
	float density = 0.0f;
	f3 g = mk3(0.0f, 0.0f, 0.0f);
	uint32_t nn = 0, cnt = 0;
	float const hh = f.kernel.h_squared;
	for (uint32_t i0 = 0; i0 < total; i0 += 32u)
	{
		uint32_t const i = i0 + (uint32_t)lane;
		bool const act = i < total;
		uint32_t j = 0;
#pragma unroll
		for (int r = 0; r < 9; r++)
			if (i - ro[r] < rnn[r]) j = rbb[r] + (i - ro[r]);          // (unsigned: also false for i < ro[r])
		bool in = false;
		float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, l2 = 0.0f;
		if (act)
		{
			float4 const q = __ldg(f.sorted + j);
			// CompactNSearch: d = x - xb; l2 = d0*d0 + d1*d1 + d2*d2 (left to right); l2 < r2
			d0 = subr(p.x, q.x); d1 = subr(p.y, q.y); d2 = subr(p.z, q.z);
			l2 = addr(addr(mulr(d0, d0), mulr(d1, d1)), mulr(d2, d2));
			in = l2 < hh;
		}
		uint32_t const m = __ballot_sync(FULL, in);
		uint32_t const add = (uint32_t)__popc(m);
		if (cnt + add > (uint32_t)kCoopCap) { coop_flush<FAST>(f, cw, cnt, nn, density, g); cnt = 0; }
		if (in) cw.ent[cnt + (uint32_t)__popc(m & ((1u << lane) - 1u))] = make_float4(d0, d1, d2, l2);
		cnt += add;
	}
	coop_flush<FAST>(f, cw, cnt, nn, density, g);
	density_out = density;
	grad_out = g;
	cand_out = total;
	nn_out = nn;
}


// phase B: the few rays that are left (about 0.3% at the default settings: rays that enter the fluid through a
// sparse region, and silhouette rays that graze it for up to MaxSteps samples).  One ray per warp, lanes = samples:
// 32 consecutive samples of the ray are evaluated at once -- the sample positions do not depend on the
// densities; the first sample at or above the threshold is the reference's hit; samples behind it are discarded
// and not counted.
//
// long_ray marches queue slot t with the whole warp.  s_dyn = the CTA's dynamic shared memory (the warps' lists first).
template <bool FAST, bool ANISO>
__device__ __forceinline__ void long_ray(const FrameView& f, const MarchParams& mp, uint32_t t, const RayQueues& rq, unsigned char* s_dyn,
										 uint32_t occ_s, LaneCounters& lc, float4* __restrict__ pos_out, float4* __restrict__ nrm_out,
										 uchar4* __restrict__ rgba_out)
{
	constexpr uint32_t FULL = 0xffffffffu;
	int const lane = threadIdx.x & 31;
	uint32_t* const list = reinterpret_cast<uint32_t*>(s_dyn) + (ANISO ? 0 : (threadIdx.x >> 5) * kListWords + lane);
	float4 const a = __ldcg(rq.q1 + 2 * (size_t)t), b = __ldcg(rq.q1 + 2 * (size_t)t + 1);
	f3 cur = mk3(a.x, a.y, a.z);
	f3 const rstep = mk3(b.x, b.y, b.z);
	StepInfo const rsi = step_info(rstep);
	uint32_t const index = __float_as_uint(a.w);
	int ri = __float_as_int(b.w);
	float4 P = make_float4(0.0f, 0.0f, 0.0f, 0.0f), N = P;
#ifdef FM_LONG_PROFILE
	// profiling build: per-ray cycle counts land in the spare control words (read back through fr_get_counters)
	long long const prof_t0 = clock64();
	long long prof_walk = 0, prof_eval = 0;
	uint32_t prof_windows = 0, prof_skips = 0, prof_steps = 0, prof_general = 0;
#endif
	for (;;)
	{
#ifdef FM_LONG_PROFILE
		long long const prof_w0 = clock64();
#endif
		// The window's sample positions.  The reference's walk is a serial chain -- position += step, look the cell up,
		// skip it if it is empty, ... -- and as such the critical path of this kernel (r02o / r03: ~420 cycles per
		// sample, ~470 per skipped cell; the slowest ray of C2 walks 67 samples and 48 empty cells in 26 us).  But as long
		// as no cell is skipped the positions are just cur + step + step + ...: a RUN of samples in occupied cells is
		// settled by all lanes at once -- lane L adds the step (L - n_valid + 1) times, one rounding per addition as the
		// reference, and looks its own cell up -- and only the first sample of the run that lands in an empty cell or
		// outside the grid goes through `advance` (every lane walks it, uniform control flow), after which the next
		// run starts from where it ended.  Rays that skip at nearly every sample (a run that settles nothing) take the
		// next few samples through `advance` directly.  C2: k_march_long 0.051 -> 0.038 ms, C3 0.073 -> 0.063 ms, C1 0.040 ->
		// 0.026 ms (r03l).  Not for the anisotropic march: its 32 000 queued rays keep every warp busy evaluating, the
		// walk is not what the kernel waits for, and the runs' extra instructions cost 8 % (2.68 -> 2.90 ms).
		f3 my_pos = cur, my_prev = cur, prv = cur;
		uint32_t skips = 0, my_skips = 0;
		int const limit = min(32, mp.max_steps - ri);
		int n_valid = 0, serial = 0;
		bool gone = false;
		while (n_valid < limit)
		{
			if ((!ANISO || FM_ANISO_SPEC_WALK) && serial == 0)
			{
				f3 q = cur, qp = cur;
				for (int k = n_valid; k < limit; k++)
					if (lane >= k) { qp = q; q = add3(q, rstep); }
				bool const mine = lane >= n_valid && lane < limit;
				bool const settled = mine && in_occupied_cell(f, q, occ_s);
				uint32_t const open = __ballot_sync(FULL, mine && !settled);
				int const b = open ? __ffs(open) - 1 : limit;         // the first sample the run does not settle
				if (mine && lane < b) { my_pos = q; my_prev = qp; my_skips = skips; }
				if (b > n_valid)
					cur = mk3(__shfl_sync(FULL, q.x, b - 1), __shfl_sync(FULL, q.y, b - 1), __shfl_sync(FULL, q.z, b - 1));
				else serial = 3;
				n_valid = b;
				if (n_valid >= limit) break;
			}
			else if (!ANISO || FM_ANISO_SPEC_WALK) serial--;
			if (advance(f, mp, rstep, rsi, cur, prv, skips, occ_s)) { gone = true; break; }
			if (lane == n_valid) { my_pos = cur; my_prev = prv; my_skips = skips; }
			n_valid++;
		}
		LaneCounters tc = {};
		SampleState<ANISO> st;
#ifdef FM_LONG_PROFILE
		long long const prof_e0 = clock64();
		prof_walk += prof_e0 - prof_w0;
		prof_windows++; prof_skips += skips & 0xffffu; prof_steps += (uint32_t)n_valid; prof_general += skips >> 16;
#endif
		float density;
		if constexpr (ANISO)
		{
#ifdef FM_NO_FMAD
			// the window's samples lie on one segment: one candidate list for the warp (aniso_list_build)
			uint32_t* const wl = reinterpret_cast<uint32_t*>(s_dyn) + (threadIdx.x >> 5) * kAnisoListCap;
			uint32_t wn;
			bool listed;
			st.have_grad = false;
			density = aniso_density_warp(f, mp, my_pos, lane < n_valid, st.as, tc, wl, wn, listed);
#else
			density = 0.0f;
#endif
		}
		else density = lane < n_valid ? sample_density<ANISO, false, false>(f, mp, my_pos, st, tc, list) : 0.0f;
		uint32_t const hits = __ballot_sync(FULL, lane < n_valid && density >= mp.iso);
#ifdef FM_LONG_PROFILE
		prof_eval += clock64() - prof_e0;
#endif
		int const kstar = hits ? __ffs(hits) - 1 : n_valid - 1;     // last sample that counts
		if (lane <= kstar) { add_counts(lc, tc); lc.steps++; }
		if (hits)
		{
			bool coop = false;
			if constexpr (!ANISO)
			{
				if (mp.bisection_steps == 0)
				{
					// the normal of the hit: gradient sum at the hit sample, the whole warp on it (the warp's list area is free)
					coop = true;
					f3 const hp = mk3(__shfl_sync(FULL, my_pos.x, kstar), __shfl_sync(FULL, my_pos.y, kstar), __shfl_sync(FULL, my_pos.z, kstar));
					CoopWarp& cw = *reinterpret_cast<CoopWarp*>(s_dyn + (threadIdx.x >> 5) * (size_t)kListWords * 4);
					float rho;
					f3 g;
					uint32_t cand, nn;
					__syncwarp();
					coop_eval<FAST>(f, hp, cw, rho, g, cand, nn);
					if (lane == kstar)
					{
						lc.skips += my_skips;
						lc.candidates += cand; lc.neighbours += nn;
						if (nn > (uint32_t)kMaxNeighbors) lc.overflow++;
						f3 const n = normalize3(g);                                   // glm::normalize(normal) (RayMarcher.cpp:338)
						P = make_float4(hp.x, hp.y, hp.z, 1.0f);
						N = make_float4(n.x, n.y, n.z, 1.0f);
						lc.hits++;
					}
				}
			}
			if (!coop && lane == kstar)
			{
				lc.skips += my_skips;
				finish_hit<ANISO, FAST>(f, mp, my_prev, my_pos, st, lc, list, P, N);
			}
			P.x = __shfl_sync(FULL, P.x, kstar); P.y = __shfl_sync(FULL, P.y, kstar);
			P.z = __shfl_sync(FULL, P.z, kstar); P.w = __shfl_sync(FULL, P.w, kstar);
			N.x = __shfl_sync(FULL, N.x, kstar); N.y = __shfl_sync(FULL, N.y, kstar);
			N.z = __shfl_sync(FULL, N.z, kstar); N.w = __shfl_sync(FULL, N.w, kstar);
			break;
		}
		if (lane == 0) { lc.skips += skips; if (gone) lc.early_exits++; }
		ri += n_valid;
		if (gone || ri >= mp.max_steps) break;
	}
	if (lane == 0) write_pixel(mp, index, P, N, pos_out, nrm_out, rgba_out);
#ifdef FM_LONG_PROFILE
	if (lane == 0)
	{
		// the ray with the longest walk: cycles << 32 | samples walked << 22 | skips << 10 | general-path skips
		unsigned long long const pa = ((unsigned long long)(uint32_t)prof_walk << 32) | ((unsigned long long)(prof_steps & 0x3ffu) << 22) |
			((unsigned long long)(prof_skips & 0xfffu) << 10) | (unsigned long long)(prof_general & 0x3ffu);
		atomicMax(reinterpret_cast<unsigned long long*>(rq.prof), pa);
		// the slowest ray: cycles << 32 | cycles evaluating
		unsigned long long const pb = ((unsigned long long)(uint32_t)(clock64() - prof_t0) << 32) | (uint32_t)prof_eval;
		atomicMax(reinterpret_cast<unsigned long long*>(rq.prof + 2), pb);
	}
#endif
	__syncwarp();
}

template <bool FAST, bool ANISO>
__global__ void __launch_bounds__(256, ANISO ? FM_ANISO_MINBLOCKS : FM_LONG_MINBLOCKS) k_march_long(FrameView f, MarchParams mp, float4* __restrict__ pos_out,
														float4* __restrict__ nrm_out, uchar4* __restrict__ rgba_out,
														RayQueues rq, DeviceCounters* __restrict__ counters, uint32_t occ_words)
{
	pdl_enter();
	constexpr uint32_t FULL = 0xffffffffu;
	int const lane = threadIdx.x & 31;
	extern __shared__ __align__(16) unsigned char s_dyn[];       // the warps' lists (kLongSmem / kAnisoSmem), then the bitmap
	LaneCounters lc = {};
	uint32_t const count = __ldcg(rq.ctl + 2);
	// the occupancy bitmap of the frame, staged once per CTA when the host found room for it (see advance)
	uint32_t occ_s = 0u;
	__shared__ __align__(8) unsigned long long s_bitmap_bar;
	bool bitmap_pending = false;
	if (occ_words != 0u && count != 0u)
	{
		uint32_t* const occ = reinterpret_cast<uint32_t*>(s_dyn + (ANISO ? kAnisoSmem : kLongSmem));
		bitmap_stage_begin(occ, f.occ_bits, occ_words, &s_bitmap_bar);        // (cudaMalloc'ed: 256-byte aligned)
		bitmap_pending = true;
		occ_s = smem_addr(occ);
	}
	uint32_t const nwarps = (gridDim.x * blockDim.x) >> 5;
	bool first = true;
	for (;;)
	{
		uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // first ray: the warp's own number (see k_march_first)
		if (!first)
		{
			if (lane == 0) t = nwarps + atomicAdd(rq.ctl + 3, 1u);
			t = __shfl_sync(FULL, t, 0);
		}
		first = false;
		if (t >= count) break;
		if (bitmap_pending) { bitmap_stage_wait(&s_bitmap_bar); bitmap_pending = false; }
		long_ray<FAST, ANISO>(f, mp, t, rq, s_dyn, occ_s, lc, pos_out, nrm_out, rgba_out);
	}
	flush_counters(lc, counters);
}


}  // namespace

}  // namespace fm
