// fm_query.cu -- point queries against the built frame, for parity checks of the neighbour search
// (Dataset::GetNeighbors, src/app/Dataset.cpp:272-280 -> NeighborhoodSearch::find_neighbors) and of the
// density / gradient sums (RayMarcher.cpp:322-336).  One thread per query point; not a hot path.
#include "fm_internal.h"

namespace fm
{

namespace
{

// the search to query: Frame::m_Search (r = h) or Frame::m_SearchExt (r = h_ext)
struct SearchView
{
	const float4* sorted;
	const uint32_t* cell_start;
	int3 kmin, kdim;
	float search_inv, r2;
};

__global__ void __launch_bounds__(128) k_query_neighbors(SearchView f, const float* __restrict__ pts, uint32_t m,
														 uint32_t* __restrict__ counts, uint32_t* __restrict__ ids, uint32_t cap)
{
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	float const px = pts[3ull * i], py = pts[3ull * i + 1], pz = pts[3ull * i + 2];
	int const big = 1 << 29;
	int const kx = min(max(search_cell_of(f.search_inv, px) - f.kmin.x, -big), big);
	int const ky = min(max(search_cell_of(f.search_inv, py) - f.kmin.y, -big), big);
	int const kz = min(max(search_cell_of(f.search_inv, pz) - f.kmin.z, -big), big);
	int const z0 = max(kz - 1, 0), z1 = min(kz + 1, f.kdim.z - 1);
	uint32_t nn = 0;
	if (z0 <= z1)
		for (int dx = -1; dx <= 1; dx++)
		{
			int const x = kx + dx;
			if ((unsigned)x >= (unsigned)f.kdim.x) continue;
			for (int dy = -1; dy <= 1; dy++)
			{
				int const y = ky + dy;
				if ((unsigned)y >= (unsigned)f.kdim.y) continue;
				uint32_t const base = ((uint32_t)x * (uint32_t)f.kdim.y + (uint32_t)y) * (uint32_t)f.kdim.z;
				uint32_t const b = f.cell_start[base + z0], e = f.cell_start[base + z1 + 1];
				for (uint32_t j = b; j < e; j++)
				{
					float4 const q = f.sorted[j];
					float const d0 = subr(px, q.x), d1 = subr(py, q.y), d2 = subr(pz, q.z);
					float const l2 = addr(addr(mulr(d0, d0), mulr(d1, d1)), mulr(d2, d2));
					if (l2 < f.r2)
					{
						if (ids && nn < cap) ids[(size_t)i * cap + nn] = __float_as_uint(q.w);
						nn++;
					}
				}
			}
		}
	counts[i] = nn;
}

__global__ void __launch_bounds__(128) k_query_density(FrameView f, const float* __restrict__ pts, uint32_t m,
													   float* __restrict__ density, float* __restrict__ grad)
{
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	float const px = pts[3ull * i], py = pts[3ull * i + 1], pz = pts[3ull * i + 2];
	int const kx = search_cell_of(f.search_inv, px) - f.kmin.x;
	int const ky = search_cell_of(f.search_inv, py) - f.kmin.y;
	int const kz = search_cell_of(f.search_inv, pz) - f.kmin.z;
	int const z0 = max(kz - 1, 0), z1 = min(kz + 1, f.kdim.z - 1);
	float rho = 0.0f;
	f3 g = mk3(0.0f, 0.0f, 0.0f);
	uint32_t nn = 0;
	if (z0 <= z1)
		for (int dx = -1; dx <= 1; dx++)
		{
			int const x = kx + dx;
			if ((unsigned)x >= (unsigned)f.kdim.x) continue;
			for (int dy = -1; dy <= 1; dy++)
			{
				int const y = ky + dy;
				if ((unsigned)y >= (unsigned)f.kdim.y) continue;
				uint32_t const base = ((uint32_t)x * (uint32_t)f.kdim.y + (uint32_t)y) * (uint32_t)f.kdim.z;
				uint32_t const b = f.cell_start[base + z0], e = f.cell_start[base + z1 + 1];
				for (uint32_t j = b; j < e; j++)
				{
					float4 const q = f.sorted[j];
					float const d0 = subr(px, q.x), d1 = subr(py, q.y), d2 = subr(pz, q.z);
					float const l2 = addr(addr(mulr(d0, d0), mulr(d1, d1)), mulr(d2, d2));
					if (l2 < f.kernel.h_squared)
					{
						// `NumNeighbors = min(neighbors.size(), MAX_NEIGHBORS)` (RayMarcher.cpp:14, :312): the sums stop there
						if (nn < 8192u)
						{
							rho = addr(rho, spline_W_inrange(f.kernel, l2));
							g = add3(g, spline_gradW_inrange(f.kernel, mk3(-d0, -d1, -d2), l2));
						}
						nn++;
					}
				}
			}
		}
	density[i] = rho;
	if (grad) { grad[3ull * i] = g.x; grad[3ull * i + 1] = g.y; grad[3ull * i + 2] = g.z; }
}

struct DevBuf
{
	void* p = nullptr;
	~DevBuf() { if (p) cudaFree(p); }
};

// self-test of the arithmetic shortcuts that claim bit-identity with the plain IEEE operations
// (fm_common.cuh: divs3_shared, rcpr): pseudo-random operands over the whole magnitude range the guard admits and
// beyond it, every result compared bit for bit with __fdiv_rn.
__device__ __forceinline__ uint32_t mix32(uint64_t& st)
{
	st = st * 6364136223846793005ull + 1442695040888963407ull;
	uint32_t x = (uint32_t)(st >> 32);
	x ^= x >> 15; x *= 0x2c1b3c6du; x ^= x >> 12; x *= 0x297a2d39u; x ^= x >> 15;
	return x;
}

__device__ __forceinline__ float random_operand(uint64_t& st, int mode)
{
	uint32_t const m = mix32(st);
	uint32_t const sign = m & 0x80000000u;
	uint32_t const frac = m & 0x007fffffu;
	uint32_t e;
	if (mode == 0) e = 127u - (mix32(st) % 30u);               // the march's range: |x| in (1e-9, 1]
	else if (mode == 1) e = 127u - 64u + (mix32(st) % 128u);   // around and across the guard (2^-60 .. 2^60)
	else e = mix32(st) % 256u;                                 // anything: zeros, denormals, infinities, NaNs
	return __uint_as_float(sign | (e << 23) | frac);
}

__global__ void __launch_bounds__(256) k_selftest_division(uint64_t per_thread, uint64_t seed, unsigned long long* mismatches)
{
	uint64_t st = seed + (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9e3779b97f4a7c15ull;
	unsigned long long bad = 0;
	for (uint64_t it = 0; it < per_thread; it++)
	{
		int const mode = (int)(it % 3);
		f3 const a = mk3(random_operand(st, mode), random_operand(st, mode), random_operand(st, mode));
		float const d = random_operand(st, mode);
		f3 const q = divs3_shared(a, d);
		f3 const w = mk3(divr(a.x, d), divr(a.y, d), divr(a.z, d));
		// NaNs compare by class only (the payload of a generated NaN is unspecified either way)
		bad += (__float_as_uint(q.x) != __float_as_uint(w.x) && !(isnan(q.x) && isnan(w.x))) ? 1 : 0;
		bad += (__float_as_uint(q.y) != __float_as_uint(w.y) && !(isnan(q.y) && isnan(w.y))) ? 1 : 0;
		bad += (__float_as_uint(q.z) != __float_as_uint(w.z) && !(isnan(q.z) && isnan(w.z))) ? 1 : 0;
		float qa, qb;
		div2_shared(a.x, a.y, d, qa, qb);
		bad += (__float_as_uint(qa) != __float_as_uint(w.x) && !(isnan(qa) && isnan(w.x))) ? 1 : 0;
		bad += (__float_as_uint(qb) != __float_as_uint(w.y) && !(isnan(qb) && isnan(w.y))) ? 1 : 0;
		float const r1 = rcpr(d), r2 = divr(1.0f, d);
		bad += (__float_as_uint(r1) != __float_as_uint(r2) && !(isnan(r1) && isnan(r2))) ? 1 : 0;
		// the spline's one-guard evaluation against the separately guarded functions it replaces (finite, non-zero r:
		// components of every magnitude, kernel radii from 2^-25 to 2^25)
		if (mode != 2)
		{
			SplineKernel k;
			k.h = __uint_as_float(((127u - 25u + (mix32(st) % 51u)) << 23) | (mix32(st) & 0x007fffffu));
			k.h_squared = mulr(k.h, k.h); k.h_inv = divr(1.0f, k.h); k.sig_d = 2.5464790894703255f;
			float const rn = addr(addr(mulr(a.x, a.x), mulr(a.y, a.y)), mulr(a.z, a.z));
			if (rn > 0.0f && rn < 0x1p120f)
			{
				float W;
				f3 g;
				spline_W_gradW_inrange(k, a, rn, W, g);
				float const W2 = spline_W_inrange(k, rn);
				f3 const g2 = spline_gradW_inrange(k, a, rn);
				bad += (__float_as_uint(W) != __float_as_uint(W2) && !(isnan(W) && isnan(W2))) ? 1 : 0;
				bad += (__float_as_uint(g.x) != __float_as_uint(g2.x) && !(isnan(g.x) && isnan(g2.x))) ? 1 : 0;
				bad += (__float_as_uint(g.y) != __float_as_uint(g2.y) && !(isnan(g.y) && isnan(g2.y))) ? 1 : 0;
				bad += (__float_as_uint(g.z) != __float_as_uint(g2.z) && !(isnan(g.z) && isnan(g2.z))) ? 1 : 0;
			}
		}
	}
	if (bad) atomicAdd(mismatches, bad);
}

// every float in [2^-100, 2^100]: sqrt_rn_normal against __fsqrt_rn, rcp_rn_normal against __frcp_rn -- for the
// reciprocal also the negative ones (the spline only takes it of lengths, but the function does not care)
__global__ void __launch_bounds__(256) k_selftest_sqrt_rcp(unsigned long long* mismatches)
{
	uint32_t const lo = 0x0d800000u, hi = 0x71800000u;        // 2^-100 .. 2^100
	unsigned long long bad = 0;
	for (uint64_t b = (uint64_t)lo + blockIdx.x * blockDim.x + threadIdx.x; b <= (uint64_t)hi; b += (uint64_t)gridDim.x * blockDim.x)
	{
		float const x = __uint_as_float((uint32_t)b);
		bad += __float_as_uint(sqrt_rn_normal(x)) != __float_as_uint(__fsqrt_rn(x)) ? 1 : 0;
		bad += __float_as_uint(rcp_rn_normal(x)) != __float_as_uint(__frcp_rn(x)) ? 1 : 0;
		bad += __float_as_uint(rcp_rn_normal(-x)) != __float_as_uint(__frcp_rn(-x)) ? 1 : 0;
	}
	if (bad) atomicAdd(mismatches, bad);
}

}  // namespace

// read-only streaming over a buffer that fits the L2 but not the L1s: the L2 -> SM bandwidth the march's gathers are
// served at (SURVEY 8d asks for this number next to the HBM figure; MEASURED_PEAKS.json has no L2 entry)
__global__ void __launch_bounds__(256) k_l2_stream(const uint4* __restrict__ buf, uint32_t n16, uint32_t reps, uint32_t* __restrict__ sink)
{
	uint32_t acc = 0;
	uint32_t const gsize = gridDim.x * blockDim.x;
	for (uint32_t r = 0; r < reps; r++)
		for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gsize)
		{
			uint4 const v = __ldcg(buf + i);      // .cg: served by the L2, not an L1
			acc += v.x ^ v.y ^ v.z ^ v.w;
		}
	if (acc == 0x9e3779b9u) *sink = acc;       // keeps the loads alive
}

int measure_l2_bandwidth(Context* ctx, size_t bytes, uint32_t reps, float* gbs)
{
	cudaStream_t const s = ctx->stream;
	DevBuf db, ds;
	size_t const n16 = bytes / 16;
	if (n16 == 0 || n16 > 0x7fffffffull) { set_error("fr_measure_l2_bandwidth: bad size"); return FR_ERR_INVALID; }
	FM_CUDA(cudaMalloc(&db.p, n16 * 16));
	FM_CUDA(cudaMalloc(&ds.p, 4));
	FM_CUDA(cudaMemsetAsync(db.p, 0x5a, n16 * 16, s));
	unsigned const blocks = (unsigned)ctx->sm_count * 8;
	k_l2_stream<<<blocks, 256, 0, s>>>((const uint4*)db.p, (uint32_t)n16, 2, (uint32_t*)ds.p);      // warm the L2
	cudaEvent_t e0, e1;
	FM_CUDA(cudaEventCreate(&e0));
	FM_CUDA(cudaEventCreate(&e1));
	FM_CUDA(cudaEventRecord(e0, s));
	k_l2_stream<<<blocks, 256, 0, s>>>((const uint4*)db.p, (uint32_t)n16, reps, (uint32_t*)ds.p);
	FM_CUDA(cudaEventRecord(e1, s));
	ctx->kernel_launches += 2;
	FM_CUDA(cudaGetLastError());
	FM_CUDA(cudaStreamSynchronize(s));
	float ms = 0.0f;
	FM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	*gbs = ms > 0.0f ? (float)((double)n16 * 16.0 * reps / (ms * 1e-3) / 1e9) : 0.0f;
	return FR_OK;
}

int selftest_division(Context* ctx, uint64_t n, uint64_t seed, uint64_t* mismatches)
{
	cudaStream_t const s = ctx->stream;
	DevBuf dm;
	FM_CUDA(cudaMalloc(&dm.p, 8));
	FM_CUDA(cudaMemsetAsync(dm.p, 0, 8, s));
	unsigned const blocks = (unsigned)ctx->sm_count * 8;
	uint64_t const per_thread = (n + (uint64_t)blocks * 256 - 1) / ((uint64_t)blocks * 256);
	k_selftest_division<<<blocks, 256, 0, s>>>(per_thread, seed, (unsigned long long*)dm.p);
	k_selftest_sqrt_rcp<<<blocks, 256, 0, s>>>((unsigned long long*)dm.p);
	ctx->kernel_launches += 2;
	FM_CUDA(cudaGetLastError());
	unsigned long long h = 0;
	FM_CUDA(cudaMemcpyAsync(&h, dm.p, 8, cudaMemcpyDeviceToHost, s));
	FM_CUDA(cudaStreamSynchronize(s));
	*mismatches = h;
	return FR_OK;
}

int query_neighbors(Context* ctx, const Frame& f, const float* points_host, size_t m, uint32_t* counts,
					uint32_t* ids, size_t cap, bool ext)
{
	if (m == 0) return FR_OK;
	if (m > 0x7fffffffull || cap > 0xffffffffull) { set_error("fr_query_neighbors: too many points"); return FR_ERR_INVALID; }
	cudaStream_t const s = ctx->stream;
	DevBuf dp, dc, di;
	FM_CUDA(cudaMalloc(&dp.p, m * 12));
	FM_CUDA(cudaMalloc(&dc.p, m * 4));
	if (ids && cap) FM_CUDA(cudaMalloc(&di.p, m * cap * 4));
	FM_CUDA(cudaMemcpyAsync(dp.p, points_host, m * 12, cudaMemcpyHostToDevice, s));
	FrameView const fv = make_view(f);
	SearchView sv;
	if (ext)
	{
		sv.sorted = fv.sorted_ext; sv.cell_start = fv.cell_start_ext; sv.kmin = fv.kmin_ext; sv.kdim = fv.kdim_ext;
		sv.search_inv = fv.search_inv_ext; sv.r2 = fv.h_ext_squared;
	}
	else
	{
		sv.sorted = fv.sorted; sv.cell_start = fv.cell_start; sv.kmin = fv.kmin; sv.kdim = fv.kdim;
		sv.search_inv = fv.search_inv; sv.r2 = fv.kernel.h_squared;
	}
	k_query_neighbors<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(sv, (const float*)dp.p, (uint32_t)m,
																 (uint32_t*)dc.p, (uint32_t*)di.p, (uint32_t)cap);
	ctx->kernel_launches += 1;
	FM_CUDA(cudaGetLastError());
	FM_CUDA(cudaMemcpyAsync(counts, dc.p, m * 4, cudaMemcpyDeviceToHost, s));
	if (di.p) FM_CUDA(cudaMemcpyAsync(ids, di.p, m * cap * 4, cudaMemcpyDeviceToHost, s));
	FM_CUDA(cudaStreamSynchronize(s));
	return FR_OK;
}

int query_density(Context* ctx, const Frame& f, const float* points_host, size_t m, float* density, float* grad)
{
	if (m == 0) return FR_OK;
	if (m > 0x7fffffffull) { set_error("fr_query_density: too many points"); return FR_ERR_INVALID; }
	cudaStream_t const s = ctx->stream;
	DevBuf dp, dd, dg;
	FM_CUDA(cudaMalloc(&dp.p, m * 12));
	FM_CUDA(cudaMalloc(&dd.p, m * 4));
	if (grad) FM_CUDA(cudaMalloc(&dg.p, m * 12));
	FM_CUDA(cudaMemcpyAsync(dp.p, points_host, m * 12, cudaMemcpyHostToDevice, s));
	k_query_density<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(make_view(f), (const float*)dp.p, (uint32_t)m,
															   (float*)dd.p, (float*)dg.p);
	ctx->kernel_launches += 1;
	FM_CUDA(cudaGetLastError());
	FM_CUDA(cudaMemcpyAsync(density, dd.p, m * 4, cudaMemcpyDeviceToHost, s));
	if (grad) FM_CUDA(cudaMemcpyAsync(grad, dg.p, m * 12, cudaMemcpyDeviceToHost, s));
	FM_CUDA(cudaStreamSynchronize(s));
	return FR_OK;
}

}  // namespace fm
