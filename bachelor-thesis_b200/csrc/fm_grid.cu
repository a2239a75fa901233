// fm_grid.cu -- GPU build of the per-frame acceleration structures (sm_100a).
//
// Replaces Frame::Frame (src/app/Dataset.cpp:9-24): BuildSearch (CompactNSearch hash + z-sort,
// :49-76), ComputeAABB (:78-92) and BuildDensityGrid (:94-165) with a uniform-grid counting sort:
//
//   k_aabb          particle min/max                      reads 12 B/particle
//   k_grid_params   m_Min/m_Max/dims/cell ranges          (1 thread; exact FP32 ops of the reference)
//   k_key_count     cell key + histogram (+ occupancy histogram)   reads 12 B, writes 4 B/particle
//   k_scan_*        exclusive prefix sum over the cell histogram   8 B/cell
//   k_scatter       counting-sort scatter into float4 SoA           reads 16 B, writes 16 B/particle
//   k_cell_order    particles of a cell into ascending original index (deterministic sums)   reads 16+ B, writes 16 B/particle
//   k_flags         OctreeNode::Flag bitmask                        4 B/cell
//
// All of it is HBM/L2-bound integer and copy work; nothing here is a contraction.
#include "fm_internal.h"

#include <math.h>
#include <string.h>

namespace fm
{

namespace
{

constexpr int kThreads = 256;

// order-preserving float <-> uint encoding for atomicMin/atomicMax
__device__ __forceinline__ uint32_t enc_ordered(float f)
{
	uint32_t const u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float dec_ordered(uint32_t u)
{
	return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Frame::ComputeAABB (Dataset.cpp:78-92): min/max are exact, so any reduction order gives the same bits.  Each block
// leaves its extrema in `partial` (6 floats per block; no atomics, nothing to initialise); as a side job the grid
// zeroes the two histogram buffers of the build up to their current capacity, which saves two memsets per frame.
__global__ void __launch_bounds__(kThreads) k_aabb(const float* __restrict__ xyz, uint32_t n, float* __restrict__ partial,
												   uint32_t* __restrict__ zero_a, uint32_t words_a,
												   uint32_t* __restrict__ zero_b, uint32_t words_b)
{
	uint32_t const gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
	for (uint32_t i = gtid; i < words_a; i += gsize) zero_a[i] = 0u;
	for (uint32_t i = gtid; i < words_b; i += gsize) zero_b[i] = 0u;
	float mn[3] = { INFINITY, INFINITY, INFINITY };
	float mx[3] = { -INFINITY, -INFINITY, -INFINITY };
	for (uint32_t i = gtid; i < n; i += gsize)
	{
#pragma unroll
		for (int a = 0; a < 3; a++)
		{
			float const v = __ldg(xyz + 3ull * i + a);
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
	__shared__ float s_mn[3][kThreads / 32], s_mx[3][kThreads / 32];
	int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (lane == 0)
		for (int a = 0; a < 3; a++) { s_mn[a][warp] = mn[a]; s_mx[a][warp] = mx[a]; }
	__syncthreads();
	if (threadIdx.x < 3)
	{
		int const a = threadIdx.x;
		float lo = s_mn[a][0], hi = s_mx[a][0];
		for (int w = 1; w < kThreads / 32; w++) { lo = fminf(lo, s_mn[a][w]); hi = fmaxf(hi, s_mx[a][w]); }
		partial[6 * blockIdx.x + a] = lo;
		partial[6 * blockIdx.x + 3 + a] = hi;
	}
}

// one block: the extrema over the blocks of k_aabb, then m_Min / m_Max / dims / cell ranges with the reference's exact
// FP32 expressions
__global__ void __launch_bounds__(kThreads) k_grid_params(GridParams* gp, const float* __restrict__ partial, uint32_t blocks, float h,
														  unsigned long long* occupied)
{
	float mn[3] = { INFINITY, INFINITY, INFINITY };
	float mx[3] = { -INFINITY, -INFINITY, -INFINITY };
	for (uint32_t b = threadIdx.x; b < blocks; b += kThreads)
#pragma unroll
		for (int a = 0; a < 3; a++)
		{
			mn[a] = fminf(mn[a], partial[6 * b + a]);
			mx[a] = fmaxf(mx[a], partial[6 * b + 3 + a]);
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
	__shared__ float s_mn[3][kThreads / 32], s_mx[3][kThreads / 32];
	int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (lane == 0)
		for (int a = 0; a < 3; a++) { s_mn[a][warp] = mn[a]; s_mx[a][warp] = mx[a]; }
	__syncthreads();
	if (threadIdx.x != 0) return;
	*occupied = 0ull;
	float const pad = mulr(1.0f, h);                 // padding = 1.0f * ParticleRadius (Dataset.cpp:89)
	float const cw = mulr(1.0f, h);                  // cellWidth = 1.0f * ParticleRadius (Dataset.cpp:96)
	float const search_inv = divr(1.0f, h);          // CompactNSearch: inverse cell size in Real
	gp->cell_width = cw;
	gp->inv_cell_width = divr(1.0f, cw);             // m_InvCellWidthVec = 1.0f / vec3(cellWidth) (:98)
	gp->search_inv = search_inv;
	for (int a = 0; a < 3; a++)
	{
		float lo = s_mn[a][0], hi = s_mx[a][0];
		for (int w = 1; w < kThreads / 32; w++) { lo = fminf(lo, s_mn[a][w]); hi = fmaxf(hi, s_mx[a][w]); }
		gp->raw_min[a] = enc_ordered(lo);            // kept for build_frame_ext (the r = h_ext search ranges)
		gp->raw_max[a] = enc_ordered(hi);
		float const mn_a = subr(lo, pad);
		float const mx_a = addr(hi, pad);
		gp->mn[a] = mn_a;
		gp->mx[a] = mx_a;
		gp->gdim[a] = (int32_t)ceilf(divr(subr(mx_a, mn_a), cw));   // int32_t(std::ceil(aabb / cellWidth)) (:102-104)
		int const k0 = search_cell_of(search_inv, lo);              // cell index is monotone in x
		int const k1 = search_cell_of(search_inv, hi);
		gp->kmin[a] = k0;
		gp->kdim[a] = k1 - k0 + 1;
	}
}

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
__global__ void __launch_bounds__(kThreads) k_key_count(const float* __restrict__ xyz, uint32_t n, BuildView b,
														uint32_t* __restrict__ keys, uint32_t* __restrict__ cell_count,
														uint32_t* __restrict__ grid_counts)
{
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float const x = __ldg(xyz + 3ull * i), y = __ldg(xyz + 3ull * i + 1), z = __ldg(xyz + 3ull * i + 2);
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

// ---- exclusive scan over m counts, in place: data[i] <- sum(data[0..i)), data[m] <- total ------------
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

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

// (src -> data: the per-cell counts stay in `src` for the scatter, their tile-local exclusive scan goes to `data`)
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

// adds to every tile the sum of the tiles in front of it (each block adds those few dozen sums up itself, which saves
// the single-block scan of the tile sums and its launch) and writes the grand total behind the last element
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
	// two block reductions (sum of the tiles in front, and -- last block -- of all tiles)
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

// counting-sort scatter.  cursor[] holds the per-cell counts and is consumed (atomicSub), so no
// second table is needed; the slot order inside a cell is arbitrary here and fixed by k_cell_order.
__global__ void __launch_bounds__(kThreads) k_scatter(const float* __restrict__ xyz, uint32_t n,
													  const uint32_t* __restrict__ keys,
													  const uint32_t* __restrict__ cell_start,
													  uint32_t* __restrict__ cursor, float4* __restrict__ sorted)
{
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t const key = keys[i];
	uint32_t const left = atomicSub(cursor + key, 1u);       // count .. 1
	uint32_t const pos = cell_start[key] + (left - 1u);
	float const x = __ldg(xyz + 3ull * i), y = __ldg(xyz + 3ull * i + 1), z = __ldg(xyz + 3ull * i + 2);
	sorted[pos] = make_float4(x, y, z, __uint_as_float(i));
}

// one thread per particle slot: the final place of a particle inside its cell is its rank by original index
// (number of cell mates with a smaller index), so that every FP32 sum over a cell runs in the reference's order
// (ascending point id) and results are reproducible from run to run and from GPU to GPU.  `unordered` holds the
// cell's particles in arrival order of the scatter atomics.
__global__ void __launch_bounds__(kThreads) k_cell_order(const float4* __restrict__ unordered, uint32_t n, BuildView b,
														 const uint32_t* __restrict__ cell_start, float4* __restrict__ sorted)
{
	uint32_t const s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	float4 const v = __ldg(unordered + s);
	uint32_t const id = __float_as_uint(v.w);
	uint32_t const key = search_key(b, v.x, v.y, v.z);
	uint32_t const cb = __ldg(cell_start + key), ce = __ldg(cell_start + key + 1);
	uint32_t rank = 0;
	for (uint32_t t = cb; t < ce; t++) rank += __float_as_uint(__ldg(&unordered[t].w)) < id ? 1u : 0u;
	sorted[cb + rank] = v;
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

// OctreeNode::Flag (Dataset.cpp:136-164).  The reference sums exp(-1000 * r) * NumParticles over the 27
// cells; expf(-1000 r) is exactly 0 for r >= 1 and 1 for r = 0, and adding exact zeros changes nothing,
// so N_c == float(NumParticles of the cell itself).  Flag = N_c * W0 > isoDensity with isoDensity = 1
// (BuildDensityGrid(1), Dataset.cpp:23).
__global__ void __launch_bounds__(kThreads) k_flags(const uint32_t* __restrict__ grid_counts, uint32_t cells, float W0,
													uint32_t* __restrict__ occ_bits, unsigned long long* occupied)
{
	uint32_t const c = blockIdx.x * blockDim.x + threadIdx.x;
	bool flag = false;
	if (c < cells)
	{
		float const rho = mulr((float)grid_counts[c], W0);
		flag = rho > 1.0f;
	}
	uint32_t const word = __ballot_sync(0xffffffffu, flag);
	if ((threadIdx.x & 31) == 0 && (c >> 5) < ((cells + 31u) >> 5))
	{
		occ_bits[c >> 5] = word;
		if (word) atomicAdd(occupied, (unsigned long long)__popc(word));
	}
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
	k_cell_order<<<pblocks, kThreads, 0, s>>>(ctx->d_sort_tmp, n32, b, f->d_cell_start_ext, f->d_sorted_ext);
	ctx->kernel_launches += 5;
	FM_CUDA(cudaGetLastError());
	f->ext_valid = true;
	return FR_OK;
}

int build_frame(Context* ctx, Frame* f, const float* d_xyz, size_t n, float h, float h_ext_mult)
{
	int const rc = build_frame_begin(ctx, f, d_xyz, n, h, h_ext_mult);
	return rc ? rc : build_frame_finish(ctx);
}

// first half: bounds, grid parameters (one host wait), table allocations.  Everything the second half launches is
// fixed after this, so a sequence lane can capture the rest of the frame into a CUDA graph
int build_frame_begin(Context* ctx, Frame* f, const float* d_xyz, size_t n, float h, float h_ext_mult)
{
	ctx->build.f = nullptr;
	if (n == 0 || n > 0x7fffffffull) { set_error("fr_upload_frame: particle count must be in [1, 2^31)"); return FR_ERR_INVALID; }
	if (!(h > 0.0f)) { set_error("fr_upload_frame: h must be positive"); return FR_ERR_INVALID; }
	cudaStream_t const s = ctx->stream;
	uint32_t const n32 = (uint32_t)n;
	f->valid = false;
	f->ext_valid = false;
	f->n = n;
	f->h = h;
	f->h_ext = h_ext_mult * h;

	FM_TIME(ctx, ctx->ev[2], s);
	if (!f->d_occupied) FM_CUDA(cudaMalloc((void**)&f->d_occupied, sizeof(unsigned long long)));
	int const aabb_blocks = (int)min((size_t)ctx->sm_count * 8, (n + kThreads - 1) / kThreads);
	int rc;
	if ((rc = ensure_capacity(&ctx->d_aabb_partial, &ctx->cap_aabb_partial, (size_t)6 * ctx->sm_count * 8))) return rc;
	// the histogram buffers as they are now (they are re-checked against this frame's sizes below)
	uint32_t* const zero_a = ctx->d_scan_tmp; size_t const cap_a = ctx->d_scan_tmp ? ctx->cap_scan_tmp : 0;
	uint32_t* const zero_b = f->d_grid_counts; size_t const cap_b = f->d_grid_counts ? f->cap_grid : 0;
	uint32_t const words_a = (uint32_t)(cap_a < 0xffffffffull ? cap_a : 0), words_b = (uint32_t)(cap_b < 0xffffffffull ? cap_b : 0);
	k_aabb<<<aabb_blocks, kThreads, 0, s>>>(d_xyz, n32, ctx->d_aabb_partial, zero_a, words_a, zero_b, words_b);
	k_grid_params<<<1, kThreads, 0, s>>>(ctx->d_gp, ctx->d_aabb_partial, (uint32_t)aabb_blocks, h, f->d_occupied);
	FM_CUDA(cudaMemcpyAsync(ctx->h_gp, ctx->d_gp, sizeof(GridParams), cudaMemcpyDeviceToHost, s));
	{ int const src = stream_sync(ctx); if (src) return src; }   // table sizes depend on the AABB
	f->gp = *ctx->h_gp;
	const GridParams& gp = f->gp;
	for (int a = 0; a < 3; a++)
		if (gp.gdim[a] <= 0 || gp.kdim[a] <= 0 || !isfinite(gp.mn[a]) || !isfinite(gp.mx[a]))
		{
			set_error("fr_upload_frame: degenerate particle bounds (NaN/inf positions?)");
			return FR_ERR_INVALID;
		}
	uint64_t const cells = (uint64_t)gp.kdim[0] * (uint64_t)gp.kdim[1] * (uint64_t)gp.kdim[2];
	uint64_t const gcells = (uint64_t)gp.gdim[0] * (uint64_t)gp.gdim[1] * (uint64_t)gp.gdim[2];
	if (cells >= 0x7fffff00ull || gcells >= 0x7fffff00ull)
	{
		set_error("fr_upload_frame: grid too large (extent / h exceeds 2^31 cells)");
		return FR_ERR_INVALID;
	}
	uint32_t const cells32 = (uint32_t)cells, gcells32 = (uint32_t)gcells;
	uint32_t const occ_words = (gcells32 + 31u) / 32u;
	uint32_t const tiles = (cells32 + kScanTile - 1) / kScanTile;

	if ((rc = ensure_capacity(&f->d_sorted, &f->cap_sorted, n))) return rc;
	if ((rc = ensure_capacity(&f->d_cell_start, &f->cap_cells, (size_t)cells32 + 1))) return rc;
	if ((rc = ensure_capacity(&f->d_grid_counts, &f->cap_grid, gcells32))) return rc;
	if ((rc = ensure_capacity(&f->d_occ_bits, &f->cap_occ_words, occ_words))) return rc;
	if ((rc = ensure_capacity(&ctx->d_keys, &ctx->cap_keys, n))) return rc;
	if ((rc = ensure_capacity(&ctx->d_sort_tmp, &ctx->cap_sort_tmp, n))) return rc;
	// scan scratch: per-cell cursor copy + tile sums
	if ((rc = ensure_capacity(&ctx->d_scan_tmp, &ctx->cap_scan_tmp, (size_t)cells32 + tiles + 2))) return rc;
	uint32_t* const d_cursor = ctx->d_scan_tmp;
	uint32_t* const d_tile_sums = ctx->d_scan_tmp + cells32;

	// k_aabb has zeroed the buffers it was given; one that had to grow (or did not exist yet) is zeroed here
	if (ctx->d_scan_tmp != zero_a || ctx->cap_scan_tmp != cap_a || words_a == 0u) FM_CUDA(cudaMemsetAsync(d_cursor, 0, (size_t)cells32 * 4, s));
	if (f->d_grid_counts != zero_b || f->cap_grid != cap_b || words_b == 0u) FM_CUDA(cudaMemsetAsync(f->d_grid_counts, 0, (size_t)gcells32 * 4, s));
	ctx->build.f = f; ctx->build.d_xyz = d_xyz; ctx->build.n32 = n32; ctx->build.cells32 = cells32; ctx->build.gcells32 = gcells32;
	ctx->build.tiles = tiles;
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
	uint32_t const n32 = ctx->build.n32, cells32 = ctx->build.cells32, gcells32 = ctx->build.gcells32, tiles = ctx->build.tiles;
	uint32_t* const d_cursor = ctx->d_scan_tmp;
	uint32_t* const d_tile_sums = ctx->d_scan_tmp + cells32;
	const GridParams& gp = f->gp;

	BuildView b;
	b.kmin = make_int3(gp.kmin[0], gp.kmin[1], gp.kmin[2]);
	b.kdim = make_int3(gp.kdim[0], gp.kdim[1], gp.kdim[2]);
	b.search_inv = gp.search_inv;
	b.mn = make_float3(gp.mn[0], gp.mn[1], gp.mn[2]);
	b.gdim = make_int3(gp.gdim[0], gp.gdim[1], gp.gdim[2]);
	b.inv_cw = gp.inv_cell_width;
	b.cw = gp.cell_width;
	b.half = 0.5f * f->h;                 // CompactNSearch: half = Real(0.5) * m_r, m_r = ParticleRadius
	b.count_mode = ctx->count_mode;

	uint32_t const pblocks = (n32 + kThreads - 1) / kThreads;
	k_key_count<<<pblocks, kThreads, 0, s>>>(d_xyz, n32, b, ctx->d_keys, d_cursor, f->d_grid_counts);
	// cell_start <- exclusive scan of the counts (counts stay in d_cursor for the scatter)
	k_scan_tiles<<<tiles, kScanThreads, 0, s>>>(d_cursor, f->d_cell_start, cells32, d_tile_sums);
	k_scan_add<<<tiles, kScanThreads, 0, s>>>(f->d_cell_start, cells32, d_tile_sums, tiles);
	k_scatter<<<pblocks, kThreads, 0, s>>>(d_xyz, n32, ctx->d_keys, f->d_cell_start, d_cursor, ctx->d_sort_tmp);
	k_cell_order<<<pblocks, kThreads, 0, s>>>(ctx->d_sort_tmp, n32, b, f->d_cell_start, f->d_sorted);
	FrameView const v = make_view(*f);
	k_flags<<<(gcells32 + kThreads - 1) / kThreads, kThreads, 0, s>>>(f->d_grid_counts, gcells32, v.kernel.sig_d,
																	  f->d_occ_bits, f->d_occupied);
	ctx->kernel_launches += 8;    // aabb, params, key_count, 2 x scan, scatter, cell_order, flags
	FM_CUDA(cudaGetLastError());
	FM_TIME(ctx, ctx->ev[3], s);
	f->valid = true;
	return FR_OK;
}

}  // namespace fm
