// fm_aniso.cu -- PerPixel_Anisotropic (src/app/AdvancedRenderer/RayMarcher.cpp:346-423) on the device: the march
// kernels of fm_march.cuh instantiated with ANISO = true, and the parity probe of WPCA / AnisotropicKernel.
// This is the only translation unit compiled with -fmad=false -DFM_NO_FMAD: fm_aniso.cuh is written as plain C
// expressions that must keep one IEEE rounding per operation (see its header comment).
#include "fm_march.cuh"

namespace fm
{

namespace
{

// one thread per query point: G (RayMarcher::WPCA over GetNeighborsExt(p)), density and gradient sums over the h-subset
__global__ void __launch_bounds__(128) k_query_aniso(FrameView f, MarchParams mp, const float* __restrict__ pts, uint32_t m,
													 float* __restrict__ density, float* __restrict__ grad, float* __restrict__ g9)
{
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	f3 const p = mk3(pts[3ull * i], pts[3ull * i + 1], pts[3ull * i + 2]);
	LaneCounters lc = {};
	AnisoSample as;
	for (int k = 0; k < 9; k++) as.G.g[k] = 0.0f;
	as.detG = 0.0f;
	float const rho = aniso_density(f, mp, p, as, lc);
	density[i] = rho;
	if (grad)
	{
		f3 const g = aniso_gradient(f, p, as);
		grad[3ull * i] = g.x; grad[3ull * i + 1] = g.y; grad[3ull * i + 2] = g.z;
	}
	if (g9)
		for (int k = 0; k < 9; k++) g9[9ull * i + k] = as.G.g[k];
}

struct DevBuf
{
	void* p = nullptr;
	~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace

int march_occupancy_aniso(int* blocks_per_sm)
{
	FM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_march_first<false, true>, 256, kAnisoFirstSmem));
	return FR_OK;
}

int launch_march_kernels_aniso(Context* ctx, const MarchLaunch& ml)
{
	cudaStream_t const st = ctx->stream;
	static bool attr_done = false;
	if (!attr_done)
	{
		FM_CUDA(cudaFuncSetAttribute(k_march_first<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
		FM_CUDA(cudaFuncSetAttribute(k_march_first<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
		FM_CUDA(cudaFuncSetAttribute(k_march_long<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
		FM_CUDA(cudaFuncSetAttribute(k_march_long<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
		attr_done = true;
	}
	if (ml.fast_normals)
		k_march_first<true, true><<<ml.ctas, 256, ml.smem_first, st>>>(ml.fv, ml.mp, ctx->d_depth, ctx->d_pos, ctx->d_nrm, ctx->d_rgba_target, ml.tiles, ml.rq, ctx->d_counters, ml.occ_words_first);
	else
		k_march_first<false, true><<<ml.ctas, 256, ml.smem_first, st>>>(ml.fv, ml.mp, ctx->d_depth, ctx->d_pos, ctx->d_nrm, ctx->d_rgba_target, ml.tiles, ml.rq, ctx->d_counters, ml.occ_words_first);
	FM_TIME(ctx, ctx->ev[11], st);
	if (ml.fast_normals)
		k_march_long<true, true><<<ml.ctas_long, 256, ml.smem_long, st>>>(ml.fv, ml.mp, ctx->d_pos, ctx->d_nrm, ctx->d_rgba_target, ml.rq, ctx->d_counters, ml.occ_words);
	else
		k_march_long<false, true><<<ml.ctas_long, 256, ml.smem_long, st>>>(ml.fv, ml.mp, ctx->d_pos, ctx->d_nrm, ctx->d_rgba_target, ml.rq, ctx->d_counters, ml.occ_words);
	FM_CUDA(cudaGetLastError());
	return FR_OK;
}

int query_aniso(Context* ctx, const Frame& f, const fr_settings& s, const float* points_host, size_t m, float* density,
				float* grad, float* g9)
{
	if (m == 0) return FR_OK;
	if (m > 0x7fffffffull) { set_error("fr_query_anisotropic: too many points"); return FR_ERR_INVALID; }
	if (!f.ext_valid) { set_error("fr_query_anisotropic: the r = h_ext search of the frame is not built"); return FR_ERR_STATE; }
	cudaStream_t const st = ctx->stream;
	DevBuf dp, dd, dg, dm;
	FM_CUDA(cudaMalloc(&dp.p, m * 12));
	FM_CUDA(cudaMalloc(&dd.p, m * 4));
	if (grad) FM_CUDA(cudaMalloc(&dg.p, m * 12));
	if (g9) FM_CUDA(cudaMalloc(&dm.p, m * 36));
	FM_CUDA(cudaMemcpyAsync(dp.p, points_host, m * 12, cudaMemcpyHostToDevice, st));
	MarchParams mp = {};
	mp.k_n = s.k_n; mp.k_r = s.k_r; mp.k_s = s.k_s; mp.n_eps = (uint32_t)s.n_eps;
	k_query_aniso<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(make_view(f), mp, (const float*)dp.p, (uint32_t)m, (float*)dd.p,
															  (float*)dg.p, (float*)dm.p);
	ctx->kernel_launches += 1;
	FM_CUDA(cudaGetLastError());
	FM_CUDA(cudaMemcpyAsync(density, dd.p, m * 4, cudaMemcpyDeviceToHost, st));
	if (grad) FM_CUDA(cudaMemcpyAsync(grad, dg.p, m * 12, cudaMemcpyDeviceToHost, st));
	if (g9) FM_CUDA(cudaMemcpyAsync(g9, dm.p, m * 36, cudaMemcpyDeviceToHost, st));
	FM_CUDA(cudaStreamSynchronize(st));
	return FR_OK;
}

}  // namespace fm
