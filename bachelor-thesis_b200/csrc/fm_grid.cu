// fm_grid.cu -- GPU build of the per-frame acceleration structures (sm_100a).
//
// Replaces Frame::Frame (src/app/Dataset.cpp:9-24): BuildSearch (CompactNSearch hash + z-sort,
// :49-76), ComputeAABB (:78-92) and BuildDensityGrid (:94-165) with a uniform-grid counting sort:
//
//   k_aabb_params   particle min/max (12 B/particle), non-finite check; the last block to finish derives m_Min/m_Max/
//                   dims/cell ranges with the reference's exact FP32 expressions, checks them against the table
//                   capacities and leaves GridParams + the march's FrameView in device memory
//   k_key_count     cell key + histogram + occupancy histogram            reads 12 B, writes 4 B/particle
//   k_scan_flags    exclusive prefix sum over the cell histogram in ONE pass (decoupled look-back, 8 B/cell)
//                   + OctreeNode::Flag bitmask (4 B/cell)
//   k_scatter       counting-sort scatter of the particle indices         reads 4 B, writes 4 B/particle
//   k_cell_order    particles of a cell into ascending original index (deterministic sums), gathers the positions
//                   into the float4 SoA                                    reads 16+ B, writes 16 B/particle
//
// Every kernel behind k_aabb_params reads the grid parameters from device memory, so a build into tables that already
// exist (every frame of a sequence but the first) does not wait for the host: the host sizes its launches by the table
// capacities, the device compares the real sizes with them and raises FM_GRID_OVERFLOW.  A copy of the parameters
// leaves for the host on a side stream as soon as k_aabb_params is done, while the rest of the build and the depth
// pre-pass keep the GPU busy; the host picks it up when it launches the march (resolve_early: the march kernels take
// the frame's geometry as kernel parameters) and rebuilds the frame with a host wait in the overflow case.
// All of it is HBM/L2-bound integer and copy work.
#include "fm_internal.h"

#include <math.h>
#include <sched.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>

namespace fm
{

namespace
{

constexpr int kThreads = 256;
constexpr int kPartialStride = 8;         // floats per block in the AABB scratch: 3 min, 3 max, non-finite flag, pad

// order-preserving float <-> uint encoding
__device__ __forceinline__ uint32_t enc_ordered(float f)
{
	uint32_t const u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Frame::ComputeAABB (Dataset.cpp:78-92) + the grid parameters.  min/max are exact, so any reduction order gives the
// reference's bits.  Each block leaves its extrema in `partial`; the block that finishes last (ticket) reduces them
// and one thread evaluates the reference's FP32 expressions for m_Min / m_Max / dims (Dataset.cpp:78-104) and the
// CompactNSearch cell ranges.  As a side job the grid zeroes the histogram / scan-state buffers of the build up to
// their current capacity, which saves the memsets.
__global__ void __launch_bounds__(kThreads) k_aabb_params(const float* __restrict__ xyz, uint32_t n, float* __restrict__ partial,
														  uint32_t* __restrict__ ticket,
														  uint32_t* __restrict__ zero_a, uint32_t words_a,
														  uint32_t* __restrict__ zero_b, uint32_t words_b,
														  float h, BuildCaps caps, GridParams* __restrict__ gp, unsigned long long* __restrict__ occupied,
														  GridParams* __restrict__ gp_host, uint32_t seq)
{
	pdl_enter();
	uint32_t const gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
	for (uint32_t i = gtid; i < words_a; i += gsize) zero_a[i] = 0u;
	for (uint32_t i = gtid; i < words_b; i += gsize) zero_b[i] = 0u;
	float mn[3] = { INFINITY, INFINITY, INFINITY };
	float mx[3] = { -INFINITY, -INFINITY, -INFINITY };
	bool bad = false;
	for (uint32_t i = gtid; i < n; i += gsize)
	{
#pragma unroll
		for (int a = 0; a < 3; a++)
		{
			float const v = __ldg(xyz + 3ull * i + a);
			bad |= !isfinite(v);                // fminf / fmaxf would silently drop a NaN
			mn[a] = fminf(mn[a], v);
			mx[a] = fmaxf(mx[a], v);
		}
	}
#pragma unroll
	for (int a = 0; a < 3; a++)
	{
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
			mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
		}
	}
	bad = __any_sync(0xffffffffu, bad);
	__shared__ float s_mn[3][kThreads / 32], s_mx[3][kThreads / 32];
	__shared__ uint32_t s_bad[kThreads / 32];
	__shared__ bool s_last;
	int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (lane == 0)
	{
		for (int a = 0; a < 3; a++) { s_mn[a][warp] = mn[a]; s_mx[a][warp] = mx[a]; }
		s_bad[warp] = bad ? 1u : 0u;
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		uint32_t any_bad = 0;
		for (int w = 0; w < kThreads / 32; w++) any_bad |= s_bad[w];
		for (int a = 0; a < 3; a++)
		{
			float lo = s_mn[a][0], hi = s_mx[a][0];
			for (int w = 1; w < kThreads / 32; w++) { lo = fminf(lo, s_mn[a][w]); hi = fmaxf(hi, s_mx[a][w]); }
			partial[kPartialStride * blockIdx.x + a] = lo;
			partial[kPartialStride * blockIdx.x + 3 + a] = hi;
		}
		partial[kPartialStride * blockIdx.x + 6] = any_bad ? 1.0f : 0.0f;
		__threadfence();
		s_last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
	}
	__syncthreads();
	if (!s_last) return;

	// ---- the last block: extrema over all blocks, then the parameters ------------------------------------------------
	__threadfence();
	for (int a = 0; a < 3; a++) { mn[a] = INFINITY; mx[a] = -INFINITY; }
	float fbad = 0.0f;
	for (uint32_t b = threadIdx.x; b < gridDim.x; b += kThreads)
	{
#pragma unroll
		for (int a = 0; a < 3; a++)
		{
			mn[a] = fminf(mn[a], __ldcg(partial + kPartialStride * b + a));
			mx[a] = fmaxf(mx[a], __ldcg(partial + kPartialStride * b + 3 + a));
		}
		fbad = fmaxf(fbad, __ldcg(partial + kPartialStride * b + 6));
	}
#pragma unroll
	for (int a = 0; a < 3; a++)
	{
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
			mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
		}
	}
	bad = __any_sync(0xffffffffu, fbad != 0.0f);
	__syncthreads();
	if (lane == 0)
	{
		for (int a = 0; a < 3; a++) { s_mn[a][warp] = mn[a]; s_mx[a][warp] = mx[a]; }
		s_bad[warp] = bad ? 1u : 0u;
	}
	__syncthreads();
	if (threadIdx.x != 0) return;
	*ticket = 0u;                                    // ready for the next build
	*occupied = 0ull;
	uint32_t status = 0;
	for (int w = 0; w < kThreads / 32; w++) status |= s_bad[w] ? (uint32_t)FM_GRID_NONFINITE : 0u;
	float const pad = mulr(1.0f, h);                 // padding = 1.0f * ParticleRadius (Dataset.cpp:89)
	float const cw = mulr(1.0f, h);                  // cellWidth = 1.0f * ParticleRadius (Dataset.cpp:96)
	float const search_inv = divr(1.0f, h);          // CompactNSearch: inverse cell size in Real
	GridParams g;
	g.cell_width = cw;
	g.inv_cell_width = divr(1.0f, cw);               // m_InvCellWidthVec = 1.0f / vec3(cellWidth) (:98)
	g.search_inv = search_inv;
	double cells = 1.0, gcells = 1.0;
	for (int a = 0; a < 3; a++)
	{
		float lo = s_mn[a][0], hi = s_mx[a][0];
		for (int w = 1; w < kThreads / 32; w++) { lo = fminf(lo, s_mn[a][w]); hi = fmaxf(hi, s_mx[a][w]); }
		g.raw_min[a] = enc_ordered(lo);              // kept for build_frame_ext (the r = h_ext search ranges)
		g.raw_max[a] = enc_ordered(hi);
		float const mn_a = subr(lo, pad);
		float const mx_a = addr(hi, pad);
		g.mn[a] = mn_a;
		g.mx[a] = mx_a;
		float const fdim = ceilf(divr(subr(mx_a, mn_a), cw));       // int32_t(std::ceil(aabb / cellWidth)) (:102-104)
		float const k0f = floorf(mulr(search_inv, lo)), k1f = floorf(mulr(search_inv, hi));
		if (!isfinite(mn_a) || !isfinite(mx_a) || !(fdim >= 1.0f) || !(fdim < 2147483000.0f) || !(fabsf(k0f) < 2147483000.0f) ||
			!(fabsf(k1f) < 2147483000.0f))
		{
			status |= (uint32_t)FM_GRID_DEGENERATE;
			g.gdim[a] = 0; g.kmin[a] = 0; g.kdim[a] = 0;
			continue;
		}
		g.gdim[a] = (int32_t)fdim;
		int const k0 = search_cell_of(search_inv, lo);              // cell index is monotone in x
		int const k1 = search_cell_of(search_inv, hi);
		g.kmin[a] = k0;
		g.kdim[a] = k1 - k0 + 1;
		cells *= (double)g.kdim[a];
		gcells *= (double)g.gdim[a];
	}
	if (!status && (cells >= 2147483000.0 || gcells >= 2147483000.0)) status |= (uint32_t)FM_GRID_DEGENERATE;
	if (!status)
	{
		uint32_t const c32 = (uint32_t)cells, g32 = (uint32_t)gcells;
		if (c32 > caps.cells || g32 > caps.gcells || (g32 + 31u) / 32u > caps.occ_words) status |= (uint32_t)FM_GRID_OVERFLOW;
	}
	g.cells = status ? 0u : (uint32_t)cells;
	g.gcells = status ? 0u : (uint32_t)gcells;
	g.status = status;
	g.max_cell = 0u;
	g.n_sorted = 0u;
	g.seq = seq;
	*gp = g;
	if (gp_host)
	{
		// the host's copy, over PCIe: everything but the sequence number, a system-wide fence, then the number the host polls
		g.seq = 0u;
		*gp_host = g;
		__threadfence_system();
		*reinterpret_cast<volatile uint32_t*>(&gp_host->seq) = seq;
	}
}

// what the particle kernels need of GridParams, loaded once per thread (every lane reads the same words: L1 broadcasts)
struct BuildView
{
	int3 kmin, kdim;
	float search_inv;
	float3 mn;
	int3 gdim;
	float inv_cw;
	float cw, half;          // cell width, half the search radius (find_neighbors_box reading FR_COUNT_CENTRE_BOX)
	int count_mode;
};

__device__ __forceinline__ BuildView load_build_view(const GridParams* __restrict__ gp, float half, int count_mode)
{
	BuildView b;
	b.kmin = make_int3(gp->kmin[0], gp->kmin[1], gp->kmin[2]);
	b.kdim = make_int3(gp->kdim[0], gp->kdim[1], gp->kdim[2]);
	b.search_inv = gp->search_inv;
	b.mn = make_float3(gp->mn[0], gp->mn[1], gp->mn[2]);
	b.gdim = make_int3(gp->gdim[0], gp->gdim[1], gp->gdim[2]);
	b.inv_cw = gp->inv_cell_width;
	b.cw = gp->cell_width;
	b.half = half;
	b.count_mode = count_mode;
	return b;
}

__device__ __forceinline__ uint32_t search_key(const BuildView& b, float x, float y, float z)
{
	int const kx = search_cell_of(b.search_inv, x) - b.kmin.x;
	int const ky = search_cell_of(b.search_inv, y) - b.kmin.y;
	int const kz = search_cell_of(b.search_inv, z) - b.kmin.z;
	return ((uint32_t)kx * (uint32_t)b.kdim.y + (uint32_t)ky) * (uint32_t)b.kdim.z + (uint32_t)kz;
}

// which of the density-grid cells c - 1, c, c + 1 of one axis hold coordinate p under the centre-box reading of
// find_neighbors_box: query point = m_Min + (float(c) + 0.5) * cellWidth (Dataset.cpp:122-123), particle counted iff
// centre - r/2 <= p < centre + r/2 with r = the search radius, all in FP32.  Rounding lets neighbouring boxes overlap
// or leave a gap by an ulp, so a coordinate can fall into two cells or into none.  Bit k = cell c - 1 + k.
__device__ __forceinline__ uint32_t centre_box_mask(float p, float mn, int c, int dim, float cw, float half)
{
	uint32_t m = 0;
#pragma unroll
	for (int k = 0; k < 3; k++)
	{
		int const cc = c - 1 + k;
		if (cc < 0 || cc >= dim) continue;
		float const centre = addr(mn, mulr(addr((float)cc, 0.5f), cw));
		if (p >= subr(centre, half) && p < addr(centre, half)) m |= 1u << k;
	}
	return m;
}

// cell key + histograms.  grid_counts = OctreeNode::NumParticles = |find_neighbors_box(cell centre)| (Dataset.cpp:117-131);
// that function exists only in the reference's un-vendored CompactNSearch fork, so its reading is a
// switch (fr_set_count_mode): FR_COUNT_CENTRE_BOX (default) = the half-open box of half-width r/2 around the query
// point, which is what the reference build under oracle/_ref does; FR_COUNT_CELL_EXACT = the node
// QueryDensityGrid(particle) returns.  The two differ only for particles within an ulp of a cell face.
constexpr uint32_t kKeyExcluded = 0xffffffffu;        // a particle the region filter left out

__global__ void __launch_bounds__(kThreads) k_key_count(const float* __restrict__ xyz, uint32_t n, const GridParams* __restrict__ gp,
														float half, int count_mode, RegionFilter rf,
														uint32_t* __restrict__ keys, uint32_t* __restrict__ cell_count,
														uint32_t* __restrict__ grid_counts)
{
	pdl_enter();
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || gp->status) return;
	BuildView const b = load_build_view(gp, half, count_mode);
	float const x = __ldg(xyz + 3ull * i), y = __ldg(xyz + 3ull * i + 1), z = __ldg(xyz + 3ull * i + 2);
	if (rf.on)
	{
		// (a conservative test: the answer only decides whether the particle is carried along, never a pixel)
		float const rx = x - rf.cam[0], ry = y - rf.cam[1], rz = z - rf.cam[2];
		bool keep = true;
#pragma unroll
		for (int k = 0; k < 4; k++) keep = keep && (rf.plane[k][0] * rx + rf.plane[k][1] * ry + rf.plane[k][2] * rz <= rf.margin);
		if (!keep) { keys[i] = kKeyExcluded; return; }
	}
	uint32_t const key = search_key(b, x, y, z);
	keys[i] = key;
	atomicAdd(cell_count + key, 1u);
	// Frame::QueryDensityGrid (Dataset.cpp:26-47)
	float const fx = floorf(mulr(subr(x, b.mn.x), b.inv_cw));
	float const fy = floorf(mulr(subr(y, b.mn.y), b.inv_cw));
	float const fz = floorf(mulr(subr(z, b.mn.z), b.inv_cw));
	if (b.count_mode == FR_COUNT_CELL_EXACT)
	{
		if (fx >= 0.0f && fx < (float)b.gdim.x && fy >= 0.0f && fy < (float)b.gdim.y && fz >= 0.0f && fz < (float)b.gdim.z)
		{
			uint32_t const c = (uint32_t)fx + (uint32_t)b.gdim.x * ((uint32_t)fy + (uint32_t)b.gdim.y * (uint32_t)fz);
			atomicAdd(grid_counts + c, 1u);
		}
		return;
	}
	// floor() of a coordinate relative to m_Min lies in [-1, dim] for every particle (m_Min = min - h); clamp for safety
	int const cx = (int)fminf(fmaxf(fx, -2.0f), (float)b.gdim.x + 1.0f), cy = (int)fminf(fmaxf(fy, -2.0f), (float)b.gdim.y + 1.0f),
		cz = (int)fminf(fmaxf(fz, -2.0f), (float)b.gdim.z + 1.0f);
	// The boxes of neighbouring cells meet at the cell faces up to a few ulps of the coordinates, so a particle that is not
	// within `eps` cells of a face -- eps = 64 ulps of the largest coordinate of the frame, in cells, at least 1e-4 -- is
	// counted for its own cell and no other without evaluating the nine boxes (99 % of the particles: 250 -> ~160
	// instructions per particle).
	{
		float const big = fmaxf(fmaxf(fabsf(b.mn.x), fabsf(b.mn.y)), fabsf(b.mn.z)) + fmaxf(fmaxf((float)b.gdim.x, (float)b.gdim.y), (float)b.gdim.z) * b.cw;
		float const eps = fmaxf(1e-4f, 64.0f * 1.1920929e-7f * big * b.inv_cw);
		float const rx = mulr(subr(x, b.mn.x), b.inv_cw) - fx, ry = mulr(subr(y, b.mn.y), b.inv_cw) - fy, rz = mulr(subr(z, b.mn.z), b.inv_cw) - fz;
		bool const interior = eps < 0.25f && rx > eps && rx < 1.0f - eps && ry > eps && ry < 1.0f - eps && rz > eps && rz < 1.0f - eps &&
			cx >= 0 && cx < b.gdim.x && cy >= 0 && cy < b.gdim.y && cz >= 0 && cz < b.gdim.z;
		if (interior)
		{
			atomicAdd(grid_counts + ((uint32_t)cx + (uint32_t)b.gdim.x * ((uint32_t)cy + (uint32_t)b.gdim.y * (uint32_t)cz)), 1u);
			return;
		}
	}
	uint32_t const mx = centre_box_mask(x, b.mn.x, cx, b.gdim.x, b.cw, b.half);
	uint32_t const my = centre_box_mask(y, b.mn.y, cy, b.gdim.y, b.cw, b.half);
	uint32_t const mz = centre_box_mask(z, b.mn.z, cz, b.gdim.z, b.cw, b.half);
	if (mx == 2u && my == 2u && mz == 2u)      // the usual case: exactly the particle's own cell
	{
		atomicAdd(grid_counts + ((uint32_t)cx + (uint32_t)b.gdim.x * ((uint32_t)cy + (uint32_t)b.gdim.y * (uint32_t)cz)), 1u);
		return;
	}
	for (int kz = 0; kz < 3; kz++)
		for (int ky = 0; ky < 3; ky++)
			for (int kx = 0; kx < 3; kx++)
				if ((mx >> kx & 1u) && (my >> ky & 1u) && (mz >> kz & 1u))
					atomicAdd(grid_counts + ((uint32_t)(cx - 1 + kx) + (uint32_t)b.gdim.x * ((uint32_t)(cy - 1 + ky) + (uint32_t)b.gdim.y * (uint32_t)(cz - 1 + kz))), 1u);
}

// ---- exclusive scan over the per-cell counts in one pass + the occupancy flags ---------------------------------------
// cell_start[i] <- sum(counts[0..i)), cell_start[cells] <- total.  Tiles of 4096 counts are handed out through a ticket
// (a block that holds ticket t knows every smaller ticket is held by a block that is running or done, so waiting for
// them cannot deadlock); a tile publishes its sum as soon as it has it, then looks back over its predecessors a warp
// at a time until it meets one that already knows its prefix (decoupled look-back).
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr unsigned long long kTileSum = 1ull << 32, kTilePrefix = 2ull << 32;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& total)
{
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		uint32_t const t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o) inc += t;
	}
	if (lane == 31) s_warp[warp] = inc;
	__syncthreads();
	if (warp == 0)
	{
		uint32_t w = lane < (kScanThreads / 32) ? s_warp[lane] : 0u;
		uint32_t winc = w;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			uint32_t const t = __shfl_up_sync(0xffffffffu, winc, o);
			if (lane >= o) winc += t;
		}
		if (lane < (kScanThreads / 32)) s_warp[lane] = winc - w;
		if (lane == 31) s_warp[32] = winc;
	}
	__syncthreads();
	total = s_warp[32];
	uint32_t const r = s_warp[warp] + inc - v;
	__syncthreads();
	return r;
}

// OctreeNode::Flag (Dataset.cpp:136-164).  The reference sums exp(-1000 * r) * NumParticles over the 27
// cells; expf(-1000 r) is exactly 0 for r >= 1 and 1 for r = 0, and adding exact zeros changes nothing,
// so N_c == float(NumParticles of the cell itself).  Flag = N_c * W0 > isoDensity with isoDensity = 1
// (BuildDensityGrid(1), Dataset.cpp:23).
__global__ void __launch_bounds__(kScanThreads) k_scan_flags(GridParams* __restrict__ gp, const uint32_t* __restrict__ counts,
															 uint32_t* __restrict__ cell_start, unsigned long long* __restrict__ tile_state,
															 uint32_t* __restrict__ ticket, const uint32_t* __restrict__ grid_counts, float W0,
															 uint32_t* __restrict__ occ_bits, unsigned long long* __restrict__ occupied)
{
	pdl_enter();
	if (gp->status) return;
	uint32_t const m = gp->cells, gcells = gp->gcells;
	uint32_t const ntiles = (m + kScanTile - 1) / kScanTile;
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__shared__ uint32_t s_warp[33];
	__shared__ uint32_t s_tile, s_prefix;
	if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
	__syncthreads();
	uint32_t const tile = s_tile;
	if (tile < ntiles)
	{
		uint32_t const base = tile * kScanTile + threadIdx.x * kScanItems;
		uint32_t v[kScanItems];
		uint32_t sum = 0, biggest = 0;
#pragma unroll
		for (int k = 0; k < kScanItems; k++)
		{
			v[k] = (base + k < m) ? counts[base + k] : 0u;
			sum += v[k];
			biggest = max(biggest, v[k]);
		}
		biggest = __reduce_max_sync(0xffffffffu, biggest);
		if (lane == 0 && biggest > kMaxCellParticles) atomicMax(&gp->max_cell, biggest);
		uint32_t total;
		uint32_t run = block_exclusive_scan(sum, s_warp, total);
		if (warp == 0)
		{
			// publish, look back
			if (lane == 0) atomicExch(tile_state + tile, (tile == 0 ? kTilePrefix : kTileSum) | total);
			uint32_t prefix = 0;
			int p = (int)tile - 1;
			while (p >= 0)
			{
				int const mine = p - lane;
				unsigned long long st = kTilePrefix;                          // lanes in front of tile 0: a zero prefix
				if (mine >= 0)
					do st = *reinterpret_cast<volatile unsigned long long*>(tile_state + mine); while ((st >> 32) == 0ull);
				uint32_t const has_prefix = __ballot_sync(0xffffffffu, (st >> 32) == 2ull);
				int const stop = has_prefix ? __ffs(has_prefix) - 1 : 32;    // nearest predecessor that knows its prefix
				uint32_t const add = lane <= stop ? (uint32_t)st : 0u;
				prefix += __reduce_add_sync(0xffffffffu, add);
				if (has_prefix) break;
				p -= 32;
			}
			if (lane == 0)
			{
				if (tile != 0) atomicExch(tile_state + tile, kTilePrefix | (unsigned long long)(prefix + total));
				s_prefix = prefix;
				if (tile == ntiles - 1) { cell_start[m] = prefix + total; gp->n_sorted = prefix + total; }
			}
		}
		__syncthreads();
		run += s_prefix;
#pragma unroll
		for (int k = 0; k < kScanItems; k++)
		{
			if (base + k < m) cell_start[base + k] = run;
			run += v[k];
		}
	}
	// occupancy flags, a warp per 32 cells
	uint32_t const warps = gridDim.x * (kScanThreads / 32);
	uint32_t found = 0;
	for (uint32_t w = blockIdx.x * (kScanThreads / 32) + warp; w * 32u < gcells; w += warps)
	{
		uint32_t const c = w * 32u + lane;
		bool flag = false;
		if (c < gcells) flag = mulr((float)grid_counts[c], W0) > 1.0f;
		uint32_t const word = __ballot_sync(0xffffffffu, flag);
		if (lane == 0) { occ_bits[w] = word; found += __popc(word); }
	}
	if (lane == 0 && found) atomicAdd(occupied, (unsigned long long)found);
}

// counting-sort scatter of the particle INDICES.  cursor[] holds the per-cell counts and is consumed (atomicSub), so
// no second table is needed; the slot order inside a cell is arbitrary here and fixed by k_cell_order.
__global__ void __launch_bounds__(kThreads) k_scatter(uint32_t n, const GridParams* __restrict__ gp,
													  const uint32_t* __restrict__ keys,
													  const uint32_t* __restrict__ cell_start,
													  uint32_t* __restrict__ cursor, uint32_t* __restrict__ slot_index)
{
	pdl_enter();
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || gp->status) return;
	uint32_t const key = keys[i];
	if (key == kKeyExcluded) return;
	uint32_t const left = atomicSub(cursor + key, 1u);       // count .. 1
	slot_index[cell_start[key] + (left - 1u)] = i;
}

// In-place ascending sort of a[0, n) by key(a[i]) with the whole CTA: the bitonic network in its normalised form --
// every comparator points the same way, the first step of a merge of size k pairs i with i ^ (k - 1), the others i with
// i ^ j -- which sorts any n: a partner beyond n is an element at +infinity that no comparator would move.
// n log^2 n / 4 compare-exchanges on global memory: the ordering of a CROWDED cell only (below).
template <typename T, typename Key>
__device__ void cta_sort_in_place(T* __restrict__ a, uint32_t n, Key key)
{
	for (uint32_t k = 2; (k >> 1) < n; k <<= 1)
	{
		for (uint32_t j = k >> 1; j > 0; j >>= 1)
		{
			uint32_t const mask = (j == (k >> 1)) ? k - 1u : j;
			__syncthreads();
			for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
			{
				uint32_t const l = i ^ mask;
				if (l > i && l < n)
				{
					T const x = a[i], y = a[l];
					if (key(x) > key(y)) { a[i] = y; a[l] = x; }
				}
			}
		}
	}
	__syncthreads();
}

// one thread per particle slot: the final place of a particle inside its cell is its rank by original index
// (number of cell mates with a smaller index), so that every FP32 sum over a cell runs in the reference's order
// (ascending point id) and results are reproducible from run to run and from GPU to GPU.  `slot_index` holds the
// cell's particles in arrival order of the scatter atomics; the positions are gathered from the input array.
// The ranking is quadratic in the cell population.  A CROWDED cell -- more than kMaxCellParticles particles: a collapsed
// simulation, or h far too large for the particle spacing -- is ordered by one CTA instead, the one that
// holds the cell's first slot: it sorts the cell's stretch of slot_index in place and gathers the positions; the
// threads of the cell's other slots do nothing.  (A crowded cell has more slots than a CTA has threads, so at most
// one starts inside a CTA.)  Frames without such a cell (k_scan_flags keeps the maximum) never reach that code.
__global__ void __launch_bounds__(kThreads) k_cell_order(const float* __restrict__ xyz, uint32_t n, GridParams* __restrict__ gp,
														 uint32_t* __restrict__ slot_index,
														 const uint32_t* __restrict__ cell_start, float4* __restrict__ sorted)
{
	pdl_enter();
	uint32_t const s = blockIdx.x * blockDim.x + threadIdx.x;
	if (gp->status) return;
	bool const crowded_frame = gp->max_cell > kMaxCellParticles;       // (uniform)
	bool const valid = s < n && s < gp->n_sorted;
	if (!crowded_frame && !valid) return;
	__shared__ uint32_t s_own[2];
	if (crowded_frame)
	{
		if (threadIdx.x == 0) s_own[0] = s_own[1] = 0u;
		__syncthreads();
	}
	if (valid)
	{
		// (non-coherent loads: only a crowded cell's stretch is ever written, and whatever a thread of such a cell reads
		// from it while its owner sorts is still one of the cell's indices, which is all it is used for)
		uint32_t const id = __ldg(slot_index + s);
		float const x = __ldg(xyz + 3ull * id), y = __ldg(xyz + 3ull * id + 1), z = __ldg(xyz + 3ull * id + 2);
		BuildView const b = load_build_view(gp, 0.0f, 0);
		uint32_t const key = search_key(b, x, y, z);
		uint32_t const cb = __ldg(cell_start + key), ce = __ldg(cell_start + key + 1);
		if (ce - cb <= kMaxCellParticles)
		{
			uint32_t rank = 0;
			for (uint32_t t = cb; t < ce; t++) rank += __ldg(slot_index + t) < id ? 1u : 0u;
			sorted[cb + rank] = make_float4(x, y, z, __uint_as_float(id));
		}
		else if (s == cb) { s_own[0] = cb; s_own[1] = ce; }
	}
	if (!crowded_frame) return;
	__syncthreads();
	uint32_t const cb = s_own[0], ce = s_own[1];
	if (ce == cb) return;                                              // (uniform) no crowded cell starts in this CTA
	cta_sort_in_place(slot_index + cb, ce - cb, [](uint32_t v) { return v; });
	for (uint32_t t = cb + threadIdx.x; t < ce; t += blockDim.x)
	{
		uint32_t const id = slot_index[t];
		sorted[t] = make_float4(__ldg(xyz + 3ull * id), __ldg(xyz + 3ull * id + 1), __ldg(xyz + 3ull * id + 2), __uint_as_float(id));
	}
}

// the r = h_ext search orders float4 records (k_scatter4): same ranking; a crowded cell's records are copied to their
// stretch of `sorted` and sorted there by original index
__global__ void __launch_bounds__(kThreads) k_cell_order4(const float4* __restrict__ unordered, uint32_t n, BuildView b,
														  const uint32_t* __restrict__ cell_start, float4* __restrict__ sorted)
{
	uint32_t const s = blockIdx.x * blockDim.x + threadIdx.x;
	__shared__ uint32_t s_own[2];
	if (threadIdx.x == 0) s_own[0] = s_own[1] = 0u;
	__syncthreads();
	if (s < n)
	{
		float4 const v = __ldg(unordered + s);
		uint32_t const id = __float_as_uint(v.w);
		uint32_t const key = search_key(b, v.x, v.y, v.z);
		uint32_t const cb = __ldg(cell_start + key), ce = __ldg(cell_start + key + 1);
		if (ce - cb <= kMaxCellParticles)
		{
			uint32_t rank = 0;
			for (uint32_t t = cb; t < ce; t++) rank += __float_as_uint(__ldg(&unordered[t].w)) < id ? 1u : 0u;
			sorted[cb + rank] = v;
		}
		else if (s == cb) { s_own[0] = cb; s_own[1] = ce; }
	}
	__syncthreads();
	uint32_t const cb = s_own[0], ce = s_own[1];
	if (ce == cb) return;
	for (uint32_t t = cb + threadIdx.x; t < ce; t += blockDim.x) sorted[t] = __ldg(unordered + t);
	cta_sort_in_place(sorted + cb, ce - cb, [](float4 v) { return __float_as_uint(v.w); });
}

// two-kernel scan of the r = h_ext build (host-sized, not on the per-frame path)
__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(const uint32_t* __restrict__ src, uint32_t* __restrict__ data, uint32_t m,
															 uint32_t* __restrict__ tile_sums)
{
	__shared__ uint32_t s_warp[33];
	uint32_t const base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
	uint32_t v[kScanItems];
	uint32_t sum = 0;
#pragma unroll
	for (int k = 0; k < kScanItems; k++)
	{
		v[k] = (base + k < m) ? src[base + k] : 0u;
		sum += v[k];
	}
	uint32_t total;
	uint32_t run = block_exclusive_scan(sum, s_warp, total);
#pragma unroll
	for (int k = 0; k < kScanItems; k++)
	{
		if (base + k < m) data[base + k] = run;
		run += v[k];
	}
	if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_add(uint32_t* __restrict__ data, uint32_t m,
														   const uint32_t* __restrict__ tile_sums, uint32_t tiles)
{
	__shared__ uint32_t s_part[kScanThreads / 32];
	__shared__ uint32_t s_off;
	uint32_t const limit = blockIdx.x == gridDim.x - 1 ? tiles : blockIdx.x;       // the last block also needs the total
	uint32_t acc_before = 0, acc_all = 0;
	for (uint32_t i = threadIdx.x; i < limit; i += kScanThreads)
	{
		uint32_t const v = tile_sums[i];
		acc_all += v;
		if (i < blockIdx.x) acc_before += v;
	}
	uint32_t total = 0;
	for (int pass = 0; pass < 2; pass++)
	{
		uint32_t v = pass == 0 ? acc_before : acc_all;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
		__syncthreads();
		if (threadIdx.x == 0)
		{
			uint32_t t = 0;
			for (int w = 0; w < kScanThreads / 32; w++) t += s_part[w];
			if (pass == 0) s_off = t; else total = t;
		}
		__syncthreads();
	}
	uint32_t const off = s_off;
	uint32_t const base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
#pragma unroll
	for (int k = 0; k < kScanItems; k++)
		if (base + k < m) data[base + k] += off;
	if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) data[m] = total;
}

// ---- second search structure (r = h_ext) from the already sorted particles ----------------------------------------
__global__ void __launch_bounds__(kThreads) k_key_count4(const float4* __restrict__ src, uint32_t n, BuildView b,
														 uint32_t* __restrict__ keys, uint32_t* __restrict__ cell_count)
{
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4 const v = __ldg(src + i);
	uint32_t const key = search_key(b, v.x, v.y, v.z);
	keys[i] = key;
	atomicAdd(cell_count + key, 1u);
}

__global__ void __launch_bounds__(kThreads) k_scatter4(const float4* __restrict__ src, uint32_t n,
													   const uint32_t* __restrict__ keys,
													   const uint32_t* __restrict__ cell_start,
													   uint32_t* __restrict__ cursor, float4* __restrict__ out)
{
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t const key = keys[i];
	uint32_t const left = atomicSub(cursor + key, 1u);
	out[cell_start[key] + (left - 1u)] = __ldg(src + i);
}

}  // namespace

FrameView make_view(const Frame& f)
{
	FrameView v;
	v.sorted = f.d_sorted;
	v.cell_start = f.d_cell_start;
	v.occ_bits = f.d_occ_bits;
	v.n = (uint32_t)f.n;
	v.kmin = make_int3(f.gp.kmin[0], f.gp.kmin[1], f.gp.kmin[2]);
	v.kdim = make_int3(f.gp.kdim[0], f.gp.kdim[1], f.gp.kdim[2]);
	v.search_inv = f.gp.search_inv;
	v.mn = make_float3(f.gp.mn[0], f.gp.mn[1], f.gp.mn[2]);
	v.mx = make_float3(f.gp.mx[0], f.gp.mx[1], f.gp.mx[2]);
	v.gdim = make_int3(f.gp.gdim[0], f.gp.gdim[1], f.gp.gdim[2]);
	v.cell_width = f.gp.cell_width;
	v.inv_cell_width = make_float3(f.gp.inv_cell_width, f.gp.inv_cell_width, f.gp.inv_cell_width);
	// CubicSplineKernel::CubicSplineKernel (Kernel.cpp:8-14), evaluated in FP32 like the reference
	float const h = f.h;
	v.kernel.h = h;
	v.kernel.h_squared = h * h;
	v.kernel.h_inv = 1.0f / h;
	volatile float t = 3.14159265358979323846264338327950288f * h;
	t = t * h;
	t = t * h;
	v.kernel.sig_d = 8.0f / t;
	v.sorted_ext = f.ext_valid ? f.d_sorted_ext : nullptr;
	v.cell_start_ext = f.ext_valid ? f.d_cell_start_ext : nullptr;
	v.kmin_ext = make_int3(f.kmin_ext[0], f.kmin_ext[1], f.kmin_ext[2]);
	v.kdim_ext = make_int3(f.kdim_ext[0], f.kdim_ext[1], f.kdim_ext[2]);
	v.search_inv_ext = f.search_inv_ext;
	v.h_ext = f.h_ext;
	v.h_ext_squared = f.h_ext * f.h_ext;
	v.aniso_sig = 8.0f / 3.14159265358979323846264338327950288f;
	return v;
}

static float host_dec_ordered(uint32_t u)
{
	uint32_t const b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
	float f;
	memcpy(&f, &b, 4);
	return f;
}

// CompactNSearch cell_index on the host (IEEE single multiply, truncation): the same bits as search_cell_of
static int host_search_cell_of(float inv, float x)
{
	volatile float const m = inv * x;
	int const t = (int)m;
	return x >= 0.0f ? t : t - 1;
}

// Frame::BuildSearch, second half (Dataset.cpp:65-75): the r = h_ext search over a copy of the particles.  Built from
// the h-sorted array (positions + original index), so the raw upload does not have to stay resident; inside a cell
// particles end up in ascending original index like in the reference's Morton-sorted copy.
int build_frame_ext(Context* ctx, Frame* f)
{
	if (f->ext_valid) return FR_OK;
	if (!f->valid) { set_error("build_frame_ext: frame not built"); return FR_ERR_STATE; }
	if (!(f->h_ext > 0.0f)) { set_error("anisotropic path: h_ext must be positive (particleRadiusMultiplier)"); return FR_ERR_INVALID; }
	cudaStream_t const s = ctx->stream;
	uint32_t const n32 = (uint32_t)f->n;
	volatile float inv_v = 1.0f / f->h_ext;
	float const inv = inv_v;
	f->search_inv_ext = inv;
	uint64_t cells = 1;
	for (int a = 0; a < 3; a++)
	{
		int const k0 = host_search_cell_of(inv, host_dec_ordered(f->gp.raw_min[a]));
		int const k1 = host_search_cell_of(inv, host_dec_ordered(f->gp.raw_max[a]));
		f->kmin_ext[a] = k0;
		f->kdim_ext[a] = k1 - k0 + 1;
		cells *= (uint64_t)f->kdim_ext[a];
	}
	if (cells >= 0x7fffff00ull) { set_error("build_frame_ext: grid too large"); return FR_ERR_INVALID; }
	uint32_t const cells32 = (uint32_t)cells;
	uint32_t const tiles = (cells32 + kScanTile - 1) / kScanTile;
	int rc;
	if ((rc = ensure_capacity(&f->d_sorted_ext, &f->cap_sorted_ext, f->n))) return rc;
	if ((rc = ensure_capacity(&f->d_cell_start_ext, &f->cap_cells_ext, (size_t)cells32 + 1))) return rc;
	if ((rc = ensure_capacity(&ctx->d_keys, &ctx->cap_keys, f->n))) return rc;
	if ((rc = ensure_capacity(&ctx->d_sort_tmp, &ctx->cap_sort_tmp, f->n))) return rc;
	if ((rc = ensure_capacity(&ctx->d_scan_tmp, &ctx->cap_scan_tmp, (size_t)cells32 + tiles + 2))) return rc;
	uint32_t* const d_cursor = ctx->d_scan_tmp;
	uint32_t* const d_tile_sums = ctx->d_scan_tmp + cells32;
	FM_CUDA(cudaMemsetAsync(d_cursor, 0, (size_t)cells32 * 4, s));
	BuildView b;
	b.kmin = make_int3(f->kmin_ext[0], f->kmin_ext[1], f->kmin_ext[2]);
	b.kdim = make_int3(f->kdim_ext[0], f->kdim_ext[1], f->kdim_ext[2]);
	b.search_inv = inv;
	b.mn = make_float3(0.0f, 0.0f, 0.0f);
	b.gdim = make_int3(0, 0, 0);
	b.inv_cw = 0.0f;
	b.cw = 0.0f; b.half = 0.0f; b.count_mode = FR_COUNT_CELL_EXACT;
	uint32_t const pblocks = (n32 + kThreads - 1) / kThreads;
	k_key_count4<<<pblocks, kThreads, 0, s>>>(f->d_sorted, n32, b, ctx->d_keys, d_cursor);
	k_scan_tiles<<<tiles, kScanThreads, 0, s>>>(d_cursor, f->d_cell_start_ext, cells32, d_tile_sums);
	k_scan_add<<<tiles, kScanThreads, 0, s>>>(f->d_cell_start_ext, cells32, d_tile_sums, tiles);
	k_scatter4<<<pblocks, kThreads, 0, s>>>(f->d_sorted, n32, ctx->d_keys, f->d_cell_start_ext, d_cursor, ctx->d_sort_tmp);
	k_cell_order4<<<pblocks, kThreads, 0, s>>>(ctx->d_sort_tmp, n32, b, f->d_cell_start_ext, f->d_sorted_ext);
	ctx->kernel_launches += 5;
	FM_CUDA(cudaGetLastError());
	f->ext_valid = true;
	return FR_OK;
}

void free_frame_small(Frame& f)
{
	if (f.d_gp) cudaFree(f.d_gp);
	if (f.h_gp) cudaFreeHost(f.h_gp);
	f.d_gp = nullptr; f.h_gp = nullptr; f.h_gp_dev = nullptr;
}

// CubicSplineKernel::sig_d = 8 / (pi h^3) (Kernel.cpp:8-14), evaluated in FP32 like the reference
static float spline_sig_d(float h)
{
	volatile float t = 3.14159265358979323846264338327950288f * h;
	t = t * h;
	t = t * h;
	return 8.0f / t;
}

// the view frustum of the context's pixel rectangle as 4 planes through the camera position (see RegionFilter)
static RegionFilter region_filter(const Context* ctx, float h)
{
	RegionFilter rf;
	memset(&rf, 0, sizeof rf);
	if (ctx->region[2] <= ctx->region[0] || !ctx->have_camera) return rf;
	const float* m = ctx->camera.inv_projection_view;
	double const W = ctx->width, H = ctx->height;
	// far-plane points of the rectangle's corners (pixel corners, as the ray generation uses them; +-1 pixel of slack)
	double c[4][3];
	double const xs[2] = { (ctx->region[0] - 1) * 2.0 / W - 1.0, (ctx->region[2] + 1) * 2.0 / W - 1.0 };
	double const ys[2] = { (ctx->region[1] - 1) * 2.0 / H - 1.0, (ctx->region[3] + 1) * 2.0 / H - 1.0 };
	int const order[4][2] = { { 0, 0 }, { 1, 0 }, { 1, 1 }, { 0, 1 } };
	for (int k = 0; k < 4; k++)
	{
		double const v[4] = { xs[order[k][0]], ys[order[k][1]], 1.0, 1.0 };
		double o[4];
		for (int r = 0; r < 4; r++) o[r] = m[0 + r] * v[0] + m[4 + r] * v[1] + m[8 + r] * v[2] + m[12 + r] * v[3];
		for (int a = 0; a < 3; a++) c[k][a] = o[a] / o[3] - ctx->camera.position[a];
	}
	double centre[3] = { 0, 0, 0 };
	for (int k = 0; k < 4; k++) for (int a = 0; a < 3; a++) centre[a] += 0.25 * c[k][a];
	for (int k = 0; k < 4; k++)
	{
		const double* a = c[k];
		const double* b = c[(k + 1) & 3];
		double nrm[3] = { a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0] };
		double len = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
		if (!(len > 0.0)) return rf;                               // degenerate camera: no filter (every particle is kept)
		double const inward = nrm[0] * centre[0] + nrm[1] * centre[1] + nrm[2] * centre[2];
		double const sgn = inward > 0.0 ? -1.0 : 1.0;              // outward normal
		for (int q = 0; q < 3; q++) rf.plane[k][q] = (float)(sgn * nrm[q] / len);
	}
	for (int a = 0; a < 3; a++) rf.cam[a] = ctx->camera.position[a];
	rf.margin = 3.6f * h;
	rf.on = 1;
	return rf;
}

// layout of ctx->d_scan_tmp for `cells` histogram entries: [cursor: cells, padded to even][tile states: 2 words per
// tile][scan ticket][AABB ticket]
static size_t scan_tmp_words(size_t cells) { return ((cells + 1) & ~(size_t)1) + 2 * ((cells + kScanTile - 1) / kScanTile + 1) + 2; }

bool build_will_not_wait(const Context* ctx, const Frame* f, bool allow_async)
{
	// tables of an earlier build of this slot that the scan scratch also covers: no host round trip
	return allow_async && ctx->async_build && f->cap_cells > 1 && f->cap_grid > 0 && f->cap_occ_words > 0 && ctx->d_scan_tmp &&
		ctx->cap_scan_tmp >= scan_tmp_words(f->cap_cells - 1) && ctx->d_aabb_partial;
}

int build_frame(Context* ctx, Frame* f, const float* d_xyz, size_t n, float h, float h_ext_mult, bool allow_async)
{
	int const rc = build_frame_begin(ctx, f, d_xyz, n, h, h_ext_mult, allow_async);
	return rc ? rc : build_frame_finish(ctx);
}

static int launch_aabb_params(Context* ctx, Frame* f, const float* d_xyz, uint32_t n32, float h, BuildCaps caps)
{
	cudaStream_t const s = ctx->stream;
	int const aabb_blocks = (int)std::min((size_t)ctx->sm_count * 8, ((size_t)n32 + kThreads - 1) / kThreads);
	// the histogram / scan-state buffers as they are now: zeroed by the same grid
	uint32_t* const zero_a = ctx->d_scan_tmp; size_t const cap_a = ctx->d_scan_tmp ? ctx->cap_scan_tmp : 0;
	uint32_t* const zero_b = f->d_grid_counts; size_t const cap_b = f->d_grid_counts ? f->cap_grid : 0;
	uint32_t const words_a = (uint32_t)(cap_a < 0xffffffffull ? cap_a : 0), words_b = (uint32_t)(cap_b < 0xffffffffull ? cap_b : 0);
	uint32_t* const ticket = (uint32_t*)(ctx->d_aabb_partial + (size_t)kPartialStride * ctx->sm_count * 8);
	FM_CUDA(launch_pdl(k_aabb_params, dim3(aabb_blocks), dim3(kThreads), 0, s, d_xyz, n32, ctx->d_aabb_partial, ticket, zero_a, words_a, zero_b, words_b, h, caps,
												  f->d_gp, f->d_occupied, f->h_gp_dev, f->gp_seq));
	ctx->kernel_launches += 1;
	FM_CUDA(cudaGetLastError());
	return FR_OK;
}

static const char* grid_status_text(uint32_t status)
{
	if (status & FM_GRID_NONFINITE) return "frame build: a particle coordinate is NaN or infinite";
	if (status & FM_GRID_DEGENERATE) return "frame build: degenerate particle bounds (extent / h exceeds 2^31 cells)";
	return "frame build failed";
}

int build_frame_begin(Context* ctx, Frame* f, const float* d_xyz, size_t n, float h, float h_ext_mult, bool allow_async)
{
	ctx->build.f = nullptr;
	if (n == 0 || n > 0x7fffffffull) { set_error("fr_upload_frame: particle count must be in [1, 2^31)"); return FR_ERR_INVALID; }
	if (!(h > 0.0f) || !isfinite(h)) { set_error("fr_upload_frame: h must be positive"); return FR_ERR_INVALID; }
	cudaStream_t const s = ctx->stream;
	uint32_t const n32 = (uint32_t)n;
	f->valid = false;
	f->ext_valid = false;
	f->gp_pending = false;
	f->gp_early_pending = false;
	f->gp_host_valid = false;
	f->n = n;
	f->h = h;
	f->h_ext = h_ext_mult * h;
	static std::atomic<uint64_t> g_build_serial{ 0 };
	f->build_serial = ++g_build_serial;        // process-wide: a Frame object may move (std::vector) and its address be reused
	f->gp_seq = (uint32_t)(f->build_serial & 0x7fffffffu) + 1u;      // never 0

	int rc;
	if (!f->d_occupied) FM_CUDA(cudaMalloc((void**)&f->d_occupied, sizeof(unsigned long long)));
	if (!f->d_gp) FM_CUDA(cudaMalloc((void**)&f->d_gp, sizeof(GridParams)));
	if (!f->h_gp)
	{
		FM_CUDA(cudaHostAlloc((void**)&f->h_gp, 2 * sizeof(GridParams), cudaHostAllocMapped));      // [0] zero-copy, [1] end of the build
		memset(f->h_gp, 0, 2 * sizeof(GridParams));
		FM_CUDA(cudaHostGetDevicePointer((void**)&f->h_gp_dev, f->h_gp, 0));
	}
	if (!ctx->d_aabb_partial)
	{
		if ((rc = ensure_capacity(&ctx->d_aabb_partial, &ctx->cap_aabb_partial, (size_t)kPartialStride * ctx->sm_count * 8 + 4))) return rc;
		FM_CUDA(cudaMemsetAsync(ctx->d_aabb_partial, 0, ctx->cap_aabb_partial * sizeof(float), s));      // the block ticket starts at 0
	}
	// everything sized by the particle count
	if ((rc = ensure_capacity(&f->d_sorted, &f->cap_sorted, n))) return rc;
	if ((rc = ensure_capacity(&ctx->d_keys, &ctx->cap_keys, n))) return rc;
	if ((rc = ensure_capacity(&ctx->d_tmp_idx, &ctx->cap_tmp_idx, n))) return rc;

	bool const async = build_will_not_wait(ctx, f, allow_async);
	ctx->build.f = f; ctx->build.d_xyz = d_xyz; ctx->build.n = n; ctx->build.h = h; ctx->build.h_ext_mult = h_ext_mult;
	ctx->build.async = async;
	f->src_xyz = d_xyz; f->src_mult = h_ext_mult;                  // for the rebuild after FM_GRID_OVERFLOW
	FM_TIME(ctx, ctx->ev[2], s);
	if (async)
	{
		ctx->build.scan_blocks = (uint32_t)((f->cap_cells - 1 + kScanTile - 1) / kScanTile);
		ctx->build.flag_cells = (uint32_t)f->cap_grid;
		return FR_OK;                   // k_aabb_params is launched by build_frame_finish with the capacities to check
	}

	BuildCaps const nocaps = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu };
	if ((rc = launch_aabb_params(ctx, f, d_xyz, n32, h, nocaps))) return rc;
	FM_CUDA(cudaMemcpyAsync(f->h_gp, f->d_gp, sizeof(GridParams), cudaMemcpyDeviceToHost, s));
	{ int const src = stream_sync(ctx); if (src) return src; }   // table sizes depend on the AABB
	f->gp = *f->h_gp;
	f->gp_host_valid = true;
	const GridParams& gp = f->gp;
	if (gp.status)
	{
		ctx->build.f = nullptr;
		set_error(grid_status_text(gp.status));
		return FR_ERR_INVALID;
	}
	uint32_t const cells32 = gp.cells, gcells32 = gp.gcells;
	uint32_t const occ_words = (gcells32 + 31u) / 32u;
	uint32_t* const zero_a = ctx->d_scan_tmp; size_t const cap_a = ctx->cap_scan_tmp;
	uint32_t* const zero_b = f->d_grid_counts; size_t const cap_b = f->cap_grid;
	if ((rc = ensure_capacity(&f->d_cell_start, &f->cap_cells, (size_t)cells32 + 1))) return rc;
	if ((rc = ensure_capacity(&f->d_grid_counts, &f->cap_grid, gcells32))) return rc;
	if ((rc = ensure_capacity(&f->d_occ_bits, &f->cap_occ_words, occ_words))) return rc;
	// scan scratch for the whole capacity of the cell table, so that later frames of this slot can build without the host
	if ((rc = ensure_capacity(&ctx->d_scan_tmp, &ctx->cap_scan_tmp, scan_tmp_words(f->cap_cells - 1)))) return rc;
	// k_aabb_params has zeroed the buffers it was given; one that had to grow (or did not exist yet) is zeroed here
	if (ctx->d_scan_tmp != zero_a || ctx->cap_scan_tmp != cap_a || !zero_a) FM_CUDA(cudaMemsetAsync(ctx->d_scan_tmp, 0, ctx->cap_scan_tmp * 4, s));
	if (f->d_grid_counts != zero_b || f->cap_grid != cap_b || !zero_b) FM_CUDA(cudaMemsetAsync(f->d_grid_counts, 0, f->cap_grid * 4, s));
	ctx->build.scan_blocks = (cells32 + kScanTile - 1) / kScanTile;
	ctx->build.flag_cells = gcells32;
	return FR_OK;
}

// second half: kernel launches only
int build_frame_finish(Context* ctx)
{
	Frame* const f = ctx->build.f;
	if (!f) { set_error("build_frame_finish without build_frame_begin"); return FR_ERR_STATE; }
	ctx->build.f = nullptr;
	cudaStream_t const s = ctx->stream;
	const float* const d_xyz = ctx->build.d_xyz;
	uint32_t const n32 = (uint32_t)ctx->build.n;
	int rc;
	size_t const cursor_cells = f->cap_cells - 1;                  // histogram capacity in cells
	// whatever is queued behind this event has the particles (the pre-pass of a render may start here: render_depth)
	if (ctx->ev_fork)
	{
		FM_CUDA(cudaEventRecord(ctx->ev_fork, s));
		ctx->fork_serial = f->build_serial;
		f->src_epoch = ctx->wait_epoch;
	}
	if (ctx->build.async)
	{
		BuildCaps caps;
		caps.cells = (uint32_t)std::min<size_t>(cursor_cells, 0x7fffffffu);
		caps.gcells = (uint32_t)std::min<size_t>(f->cap_grid, 0x7fffffffu);
		caps.occ_words = (uint32_t)std::min<size_t>(f->cap_occ_words, 0x7fffffffu);
		caps.scan_tiles = ctx->build.scan_blocks;
		if ((rc = launch_aabb_params(ctx, f, d_xyz, n32, ctx->build.h, caps))) return rc;
		f->gp_early_pending = true;          // the kernel writes them into the host's mapped copy itself (resolve_early)
	}
	uint32_t* const d_cursor = ctx->d_scan_tmp;
	size_t const state_off = (cursor_cells + 1) & ~(size_t)1;
	unsigned long long* const tile_state = (unsigned long long*)(ctx->d_scan_tmp + state_off);
	uint32_t* const scan_ticket = ctx->d_scan_tmp + state_off + 2 * ((cursor_cells + kScanTile - 1) / kScanTile + 1);

	uint32_t const pblocks = (n32 + kThreads - 1) / kThreads;
	float const half = 0.5f * f->h;                   // CompactNSearch: half = Real(0.5) * m_r, m_r = ParticleRadius
	RegionFilter const rf = region_filter(ctx, f->h);
	f->filtered = rf.on != 0;
	FM_CUDA(launch_pdl(k_key_count, dim3(pblocks), dim3(kThreads), 0, s, d_xyz, n32, f->d_gp, half, ctx->count_mode, rf, ctx->d_keys, d_cursor, f->d_grid_counts));
	// cell_start <- exclusive scan of the counts (counts stay in d_cursor for the scatter); occupancy flags
	uint32_t const flag_blocks = (ctx->build.flag_cells + kScanThreads - 1) / kScanThreads;
	uint32_t const scan_grid = std::max(ctx->build.scan_blocks, std::min(flag_blocks, (uint32_t)ctx->sm_count * 4u));
	FM_CUDA(launch_pdl(k_scan_flags, dim3(scan_grid), dim3(kScanThreads), 0, s, f->d_gp, d_cursor, f->d_cell_start, tile_state, scan_ticket, f->d_grid_counts,
													   spline_sig_d(f->h), f->d_occ_bits, f->d_occupied));
	FM_CUDA(launch_pdl(k_scatter, dim3(pblocks), dim3(kThreads), 0, s, n32, f->d_gp, ctx->d_keys, f->d_cell_start, d_cursor, ctx->d_tmp_idx));
	FM_CUDA(launch_pdl(k_cell_order, dim3(pblocks), dim3(kThreads), 0, s, d_xyz, n32, f->d_gp, ctx->d_tmp_idx, f->d_cell_start, f->d_sorted));
	ctx->kernel_launches += 4;
	FM_CUDA(cudaGetLastError());
	// the status word at the end of the build comes back with the frame's results
	FM_CUDA(cudaMemcpyAsync(f->h_gp + 1, f->d_gp, sizeof(GridParams), cudaMemcpyDeviceToHost, s));
	f->gp_pending = true;
	FM_TIME(ctx, ctx->ev[3], s);
	f->valid = true;
	return FR_OK;
}

// The host copy of the grid parameters of a frame whose build was queued without a host wait: polls the sequence
// number k_aabb_params stores, last, into mapped pinned memory -- not the stream.  FR_RETRIED: the tables of the slot
// were too small, the frame has been rebuilt (with a host wait) -- whatever was queued on it in between ran on an
// unusable frame.
int resolve_early(Context* ctx, Frame* f)
{
	if (f->gp_host_valid || !f->gp_early_pending) return FR_OK;
	volatile uint32_t* const seq = &f->h_gp[0].seq;
	for (uint32_t spins = 1; *seq != f->gp_seq; spins++)
	{
		if (ctx->blocking_sync && (spins & 15u) == 0u) sched_yield();      // oversubscribed hosts: give the core away between polls
		else __builtin_ia32_pause();
		if ((spins & 0xfffffu) == 0u)                                      // every ~million polls: has the stream failed?
		{
			cudaError_t const e = cudaStreamQuery(ctx->stream);
			if (e != cudaSuccess && e != cudaErrorNotReady) return cuda_fail(e, "cudaStreamQuery", __FILE__, __LINE__);
			if (e == cudaSuccess && *seq != f->gp_seq)
			{
				set_error("frame build: the grid parameters never arrived");
				return FR_ERR_STATE;
			}
		}
	}
	__sync_synchronize();
	f->gp_early_pending = false;
	GridParams const gp = f->h_gp[0];
	if (gp.status & FM_GRID_OVERFLOW)
	{
		// the tables of the earlier frame are too small for this one: once more, with the host sizing them
		int const rc = build_frame(ctx, f, f->src_xyz, f->n, f->h, f->src_mult, false);
		return rc ? rc : FR_RETRIED;
	}
	if (gp.status)
	{
		f->valid = false;
		f->gp_pending = false;
		set_error(grid_status_text(gp.status));
		return FR_ERR_INVALID;
	}
	f->gp = gp;
	f->gp_host_valid = true;
	return FR_OK;
}

// end of the build (the caller has drained the stream, or `synced` is false and this does): late status bits
int resolve_frame(Context* ctx, Frame* f, bool synced)
{
	if (!f->gp_pending) return FR_OK;
	int const erc = resolve_early(ctx, f);
	if (erc < 0) return erc;
	if (!synced || erc == FR_RETRIED) { int const src = stream_sync(ctx); if (src) return src; }
	f->gp_pending = false;
	GridParams const gp = f->h_gp[1];
	if (gp.status)
	{
		f->valid = false;
		set_error(grid_status_text(gp.status));
		return FR_ERR_INVALID;
	}
	return erc;
}

}  // namespace fm
