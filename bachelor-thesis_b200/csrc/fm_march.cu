// fm_march.cu -- host side of the march: parameters, work lists, kernel launches (device code: fm_march.cuh).
#include "fm_march.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace fm
{

// k_march_long's launch shape: when the frame's occupancy bitmap fits next to the warps' lists it is staged in shared
// memory (the empty-space walk is a chain of dependent bitmap lookups, see advance) and the resident CTAs per SM follow
// from what is left.  Returns the CTA count.
constexpr size_t kLongSmemMax = 200 * 1024;
static uint32_t long_launch_shape(Context* ctx, const Frame& f, const void*, size_t base_smem, int base_ctas_per_sm, uint32_t* occ_words,
								  size_t* smem, bool keep_ctas = false)
{
	*occ_words = 0;
	*smem = base_smem;
	int per_sm = base_ctas_per_sm > 0 ? base_ctas_per_sm : 1;
	static int const allow = [] { const char* e = getenv("FLUIDMARCH_OCC_SMEM"); return (e && e[0] == '0') ? 0 : 1; }();
	size_t const words = ((size_t)f.gp.gcells + 31u) / 32u;
	size_t const bytes = ((words * 4u) + 15u) & ~(size_t)15u;
	bool const fits = base_smem + bytes <= kLongSmemMax && (!keep_ctas || (base_smem + bytes + 1024u) * (size_t)per_sm <= (size_t)(228 * 1024));
	if (allow && f.gp_host_valid && words != 0 && fits)
	{
		*occ_words = (uint32_t)words;
		*smem = base_smem + bytes;
		int const fit = (int)((size_t)(228 * 1024) / (*smem + 1024u));
		if (fit < per_sm) per_sm = fit > 0 ? fit : 1;
		// one resident CTA per SM serialises the rays (r02s, C3: 69 KB of bitmap next to 48 KB of lists: 0.072 -> 0.084 ms)
		if (per_sm < 2 && base_ctas_per_sm >= 2)
		{
			*occ_words = 0; *smem = base_smem;
			per_sm = base_ctas_per_sm;
		}
	}
	return (uint32_t)(ctx->sm_count * per_sm);
}

int launch_march(Context* ctx, const Frame& f, bool do_march, bool do_shade)
{
	bool const aniso = ctx->settings.enable_anisotropy != 0;
	if (do_march && aniso && !f.ext_valid) { set_error("launch_march: the r = h_ext search of the frame is not built"); return FR_ERR_STATE; }
	MarchParams mp;
	mp.W = ctx->width; mp.H = ctx->height;
	mp.two_w_inv = 2.0f / (float)ctx->width;
	mp.two_h_inv = 2.0f / (float)ctx->height;
	mp.inv_w = 1.0f / (float)ctx->width;
	mp.inv_h = 1.0f / (float)ctx->height;
	for (int k = 0; k < 16; k++) mp.ipv[k] = ctx->camera.inv_projection_view[k];
	for (int k = 0; k < 3; k++) { mp.cam[k] = ctx->camera.position[k]; mp.dir[k] = ctx->camera.direction[k]; }
	mp.max_steps = ctx->settings.max_steps;
	mp.step_size = ctx->settings.step_size;
	mp.iso = ctx->settings.iso_density;
	mp.bisection_steps = ctx->settings.bisection_steps;
	mp.skip_last_pixel = ctx->settings.skip_last_pixel;
	// a sample outside the grid can, through rounding of m_Min = min - h, still see a particle at
	// distance h(1 - 1e-7) whose W is ~1e-18 * W0; the early exit is exact as long as that cannot
	// reach the threshold
	mp.early_out = ctx->settings.iso_density > 1e-6f ? 1 : 0;
	// the anisotropic kernel of such a grazing particle is not small (G shrinks r), so there the ray is only stopped a
	// whole cell beyond the grid, where every particle is farther than 2h - rounding
	mp.early_margin = aniso ? 1.0f : 0.0f;
	mp.part_rank = ctx->part_rank; mp.part_world = ctx->part_world;
	mp.part_tw = ctx->part_tw; mp.part_th = ctx->part_th;
	mp.part_tiles_x = (ctx->width + ctx->part_tw - 1) / ctx->part_tw;
	bool const region = ctx->region[2] > ctx->region[0];
	mp.rx0 = region ? ctx->region[0] : 0; mp.ry0 = region ? ctx->region[1] : 0;
	mp.rx1 = region ? ctx->region[2] : 0; mp.ry1 = region ? ctx->region[3] : 0;
	if (region) mp.part_world = 1;
	mp.do_march = do_march ? 1 : 0;
	mp.do_shade = do_shade ? 1 : 0;
	mp.k_n = ctx->settings.k_n; mp.k_r = ctx->settings.k_r; mp.k_s = ctx->settings.k_s;
	mp.n_eps = (uint32_t)ctx->settings.n_eps;      // `const uint32_t N_eps = m_Settings.N_eps` (RayMarcher.cpp:232)

	int const tiles_x = (ctx->width + 7) / 8, tiles_y = (ctx->height + 3) / 4;
	mp.tiles_x = tiles_x;
	static int const skip_empty = [] { const char* e = getenv("FLUIDMARCH_ANISO_EMPTY"); return (e && e[0] == '0') ? 0 : 1; }();
	mp.aniso_skip_empty = (skip_empty && ctx->settings.iso_density > 0.0f) ? 1 : 0;      // (density 0 must not be a hit)
	{
		// (read at every launch, not once: tests/test_gpu_parity.py switches it between two renders of one process)
		const char* const e = getenv("FLUIDMARCH_BGFAST");
		mp.bg_fast = (e && e[0] == '0') ? 0 : 1;
	}
	{
		const float* m = mp.ipv;
		float big = 0.0f;
		for (int r = 0; r < 3; r++)
		{
			mp.bg_c[r] = m[8 + r] + m[12 + r];
			big = std::max(big, std::fabs(m[r]) + std::fabs(m[4 + r]) + std::fabs(mp.bg_c[r]));
		}
		mp.bg_err = 64.0f * 1.1920929e-7f * big;
		mp.bg_cam = std::fabs(mp.cam[0]) + std::fabs(mp.cam[2]) + 1.0f;
	}
	size_t const npix = (size_t)ctx->width * ctx->height;
	int rc;
	if ((rc = ensure_capacity(&ctx->d_tiles, &ctx->cap_tiles, (size_t)tiles_x * tiles_y + 8))) return rc;
	if (do_march && (rc = ensure_capacity(&ctx->d_rayq, &ctx->cap_rayq, 2 * npix))) return rc;   // 32 B per pixel
	RayQueues rq;
	rq.ctl = ctx->d_counters->ctl;
	rq.prof = reinterpret_cast<uint32_t*>(&ctx->d_counters->first_examined);
	rq.q1 = ctx->d_rayq;
	uint32_t* const tiles = ctx->d_tiles;
	cudaStream_t const st = ctx->stream;
	if (!ctx->zero_counters_in_depth)
		FM_CUDA(cudaMemsetAsync(ctx->d_counters, 0, sizeof(DeviceCounters), st));      // counters + control words
	ctx->zero_counters_in_depth = false;
	dim3 const grid(((region ? mp.rx1 - mp.rx0 : ctx->width) + 31) / 32, ((region ? mp.ry1 - mp.ry0 : ctx->height) + 7) / 8);
	// The outputs of the uncovered pixels (80 % of the image at C2: 53 MB of zeros and the floor colour) do not concern
	// the march.  With stage timing off k_classify only lists the tiles, and a second launch of it writes those pixels on
	// the side stream while k_march_long runs -- a kernel that is as long as its slowest ray and leaves most of the GPU
	// idle (r03p: one C2 frame 0.2885 -> 0.2768 ms, C1 0.1096 -> 0.1008 ms, C3 unchanged; beside k_march_FIRST instead it
	// competes for the SMs: r03b, nothing gained).  FLUIDMARCH_BG=0 keeps them in the one k_classify.
	static int const bgmode = [] { const char* e = getenv("FLUIDMARCH_BG"); return e ? atoi(e) : 1; }();
	bool const beside = bgmode != 0 && do_march && !aniso && ctx->overlap_depth && !ctx->stage_timing && ctx->stream_depth != nullptr;
	FM_CUDA(launch_pdl(k_classify, dim3(grid), dim3(256), 0, st, mp, ctx->d_depth, ctx->d_pos, ctx->d_nrm, ctx->d_rgba_target, tiles, rq.ctl, beside ? 1 : 3));
	ctx->kernel_launches += 1;
	FM_TIME(ctx, ctx->ev[10], st);
	if (do_march)
	{
		// persistent: as many CTAs as stay resident (occupancy of this build), never more warps than tiles
		bool const fast = ctx->settings.fast_normals != 0;
		if (aniso)
		{
			int& per_sm = ctx->march_ctas_per_sm_aniso;
			if (per_sm == 0)
			{
				int nb = 0;
				if ((rc = march_occupancy_aniso(&nb))) return rc;
				per_sm = nb > 0 ? nb : 1;
				if (const char* e = getenv("FR_MARCH_CTAS_PER_SM")) { int const v = atoi(e); if (v > 0 && v < per_sm) per_sm = v; }   // tuning switch
			}
			uint32_t const max_ctas = (uint32_t)((tiles_x * tiles_y + 7) / 8);
			uint32_t ctas = (uint32_t)(ctx->sm_count * per_sm);
			if (ctas > max_ctas) ctas = max_ctas;
			MarchLaunch ml;
			ml.fv = make_view(f); ml.mp = mp; ml.rq = rq; ml.tiles = tiles; ml.ctas = ctas; ml.fast_normals = fast;
			ml.occ_words = 0; ml.smem_long = kAnisoSmem;
			ml.ctas_long = std::min(long_launch_shape(ctx, f, nullptr, kAnisoSmem, per_sm, &ml.occ_words, &ml.smem_long), max_ctas);
			ml.occ_words_first = 0; ml.smem_first = kAnisoFirstSmem;
			long_launch_shape(ctx, f, nullptr, kAnisoFirstSmem, per_sm, &ml.occ_words_first, &ml.smem_first, true);
			if ((rc = launch_march_kernels_aniso(ctx, ml))) return rc;
		}
		else
		{
			// k_march_first: dynamic shared memory (lists of the walk, or one stage per warp: above the 48 KB default)
			size_t const smem_first = kFirstSmem;
			int const first_warps = kFirstThreads / 32;
			FrameView const fv = make_view(f);
			auto const first = fast ? k_march_first<true, false> : k_march_first<false, false>;
			auto const longk = fast ? k_march_long<true, false> : k_march_long<false, false>;
			if (ctx->march_ctas_per_sm == 0)
			{
				FM_CUDA(cudaFuncSetAttribute(k_march_first<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLongSmemMax));
				FM_CUDA(cudaFuncSetAttribute(k_march_first<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLongSmemMax));
				int nb = 0, nl = 0;
				FM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_march_first<false, false>, kFirstThreads, smem_first));
				FM_CUDA(cudaFuncSetAttribute(k_march_long<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLongSmemMax));
				FM_CUDA(cudaFuncSetAttribute(k_march_long<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLongSmemMax));
				FM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nl, k_march_long<false, false>, 256, kLongSmem));
				if (const char* e = getenv("FR_MARCH_CTAS_PER_SM")) { int const v = atoi(e); if (v > 0 && v < nb) nb = v; }   // tuning switch
				ctx->march_ctas_per_sm = nb > 0 ? nb : 1;
				ctx->march_long_ctas_per_sm = nl > 0 ? nl : 1;
			}
			uint32_t const ntiles = (uint32_t)(tiles_x * tiles_y);
			uint32_t ctas_first = (uint32_t)(ctx->sm_count * ctx->march_ctas_per_sm);
			ctas_first = std::min(ctas_first, (ntiles + first_warps - 1) / first_warps);
			uint32_t occ_words = 0;
			size_t smem_long = kLongSmem;
			uint32_t ctas_long = long_launch_shape(ctx, f, (const void*)k_march_long<false, false>, kLongSmem, ctx->march_long_ctas_per_sm, &occ_words, &smem_long);
			ctas_long = std::min(ctas_long, (ntiles + 7) / 8);
			uint32_t occ_first = 0;
			size_t smem_first_occ = smem_first;
			long_launch_shape(ctx, f, nullptr, smem_first, ctx->march_ctas_per_sm, &occ_first, &smem_first_occ, true);
			FM_CUDA(launch_pdl(first, dim3(ctas_first), dim3(kFirstThreads), smem_first_occ, st, fv, mp, ctx->d_depth, ctx->d_pos, ctx->d_nrm, ctx->d_rgba_target, tiles, rq, ctx->d_counters, occ_first));
			FM_TIME(ctx, ctx->ev[11], st);
			if (beside)
			{
				FM_CUDA(cudaEventRecord(ctx->ev_fork2, st));
				FM_CUDA(cudaStreamWaitEvent(ctx->stream_depth, ctx->ev_fork2, 0));
				k_classify<<<grid, 256, 0, ctx->stream_depth>>>(mp, ctx->d_depth, ctx->d_pos, ctx->d_nrm, ctx->d_rgba_target, tiles, rq.ctl, 2);
				FM_CUDA(cudaEventRecord(ctx->ev_join2, ctx->stream_depth));
				ctx->kernel_launches += 1;
			}
			FM_CUDA(launch_pdl(longk, dim3(ctas_long), dim3(256), smem_long, st, fv, mp, ctx->d_pos, ctx->d_nrm, ctx->d_rgba_target, rq, ctx->d_counters, occ_words));
		}
		ctx->kernel_launches += 2;
	}
	else FM_TIME(ctx, ctx->ev[11], st);
	if (beside) FM_CUDA(cudaStreamWaitEvent(st, ctx->ev_join2, 0));
	FM_CUDA(cudaGetLastError());
	return FR_OK;
}

}  // namespace fm

#ifdef FM_FIRST_PROFILE
// profiling build only (not in include/fluidmarch.h): the per-warp records of the last k_march_first launch
extern "C" int fr_debug_first_profile(unsigned long long* out, int words)
{
	if (words > 4 * 16384) words = 4 * 16384;
	return cudaMemcpyFromSymbol(out, fm::g_first_prof, (size_t)words * 8u) == cudaSuccess ? 0 : -1;
}
#endif
