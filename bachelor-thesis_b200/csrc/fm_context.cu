// fm_context.cu -- the C ABI (include/fluidmarch.h): context, memory, stream ordering, hand-off.
//
// Host-side replacement of the reference's RayMarcher shell (src/app/AdvancedRenderer/RayMarcher.cpp:64-112:
// constructor, Prepare, Start, IsDone, Exit) and ThreadPool (src/app/ThreadPool.cpp): "Start" is a kernel
// launch on the context stream, "IsDone" a cudaEventQuery.  No CPU fallback exists: without a CUDA device
// fr_create fails with FR_ERR_NO_DEVICE.
#include "fm_internal.h"

#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

namespace fm
{

bool pdl_enabled()
{
	static int const on = [] { const char* e = getenv("FLUIDMARCH_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
	return on != 0;
}


static thread_local std::string g_error;

void set_error(const std::string& msg) { g_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
	char buf[512];
	snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
	g_error = buf;
	cudaGetLastError();   // clear the sticky-free error state
	return FR_ERR_CUDA;
}

// cuStreamWriteValue32: a stream memory operation -- executed by the front end when the stream gets there, no SM
// and no copy engine involved, so it is not held up by other lanes' persistent kernels the way a signalling kernel
// is.  Fetched through the runtime so the library does not link libcuda (it must load on machines without a driver).
typedef int (*StreamWriteValue32Fn)(cudaStream_t, unsigned long long, uint32_t, unsigned int);
static StreamWriteValue32Fn stream_write_value32()
{
	static StreamWriteValue32Fn fn = [] {
		void* p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
		{
			cudaGetLastError();
			p = nullptr;
		}
		return (StreamWriteValue32Fn)p;
	}();
	return fn;
}

int stream_sync(Context* c)
{
	c->wait_epoch++;                     // a host wait: the raw particle array of a queued build is the caller's again (render_depth)
	// the lanes of a sequence: a word in mapped pinned memory, written behind everything on the stream, polled
	// without entering the driver
	if (c->blocking_sync && c->h_sync_flag)
	{
		StreamWriteValue32Fn const write32 = stream_write_value32();
		uint32_t const seq = ++c->sync_seq;
		if (!write32 || write32(c->stream, (unsigned long long)(uintptr_t)c->d_sync_flag, seq, 0u) != 0)
		{
			FM_CUDA(cudaStreamSynchronize(c->stream));
			return FR_OK;
		}
		volatile uint32_t* const flag = c->h_sync_flag;
		uint32_t spins = 0;
		while ((int32_t)(*flag - seq) < 0)
		{
			if ((++spins & 63u) == 0u)
			{
				sched_yield();
				if ((spins & 0xfffffu) == 0u)       // every ~million polls: has the stream failed?
				{
					cudaError_t const e = cudaStreamQuery(c->stream);
					if (e != cudaSuccess && e != cudaErrorNotReady) return cuda_fail(e, "cudaStreamQuery", __FILE__, __LINE__);
				}
			}
			else __builtin_ia32_pause();
		}
		return FR_OK;
	}
	FM_CUDA(cudaStreamSynchronize(c->stream));
	return FR_OK;
}

static void free_frame(Frame& f)
{
	free_frame_small(f);
	if (f.d_sorted) cudaFree(f.d_sorted);
	if (f.d_cell_start) cudaFree(f.d_cell_start);
	if (f.d_grid_counts) cudaFree(f.d_grid_counts);
	if (f.d_occ_bits) cudaFree(f.d_occ_bits);
	if (f.d_occupied) cudaFree(f.d_occupied);
	if (f.d_sorted_ext) cudaFree(f.d_sorted_ext);
	if (f.d_cell_start_ext) cudaFree(f.d_cell_start_ext);
	f = Frame();
}

static void free_images(Context* c)
{
	if (c->d_depth) cudaFree(c->d_depth);
	if (c->d_pos_own) cudaFree(c->d_pos_own);
	if (c->d_nrm_own) cudaFree(c->d_nrm_own);
	if (c->d_rgba) cudaFree(c->d_rgba);
	c->d_depth = nullptr; c->d_pos = nullptr; c->d_nrm = nullptr; c->d_pos_own = nullptr; c->d_nrm_own = nullptr;
	c->d_rgba = nullptr; c->d_rgba_target = nullptr;
}

static int alloc_images(Context* c, int w, int h)
{
	size_t const npix = (size_t)w * (size_t)h;
	FM_CUDA(cudaMalloc((void**)&c->d_depth, npix * sizeof(float)));
	FM_CUDA(cudaMalloc((void**)&c->d_pos_own, npix * sizeof(float4)));
	FM_CUDA(cudaMalloc((void**)&c->d_nrm_own, npix * sizeof(float4)));
	c->d_pos = c->d_pos_own; c->d_nrm = c->d_nrm_own;
	FM_CUDA(cudaMalloc((void**)&c->d_rgba, npix * sizeof(uchar4)));
	FM_CUDA(cudaMemsetAsync(c->d_pos, 0, npix * sizeof(float4), c->stream));
	FM_CUDA(cudaMemsetAsync(c->d_nrm, 0, npix * sizeof(float4), c->stream));
	FM_CUDA(cudaMemsetAsync(c->d_rgba, 0, npix * sizeof(uchar4), c->stream));
	c->d_rgba_target = c->d_rgba;
	c->width = w; c->height = h;
	c->have_depth = false;
	return FR_OK;
}

static Frame* get_frame(Context* c, int frame, bool must_be_valid)
{
	if (frame < 0 || (size_t)frame >= c->frames.size() || (must_be_valid && !c->frames[frame].valid))
	{
		set_error("frame index does not name an uploaded frame");
		return nullptr;
	}
	return &c->frames[frame];
}

// upload / grid build times of the last frame build, once its events have completed
static void read_build_timings(Context* c)
{
	if (!c->stage_timing) { c->build_timed = 0; return; }
	float ms = 0.0f;
	if (c->build_timed == 2) { if (cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]) == cudaSuccess) c->timings.upload_ms = ms; }
	else if (c->build_timed == 1) c->timings.upload_ms = 0.0f;
	if (c->build_timed && cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]) == cudaSuccess) c->timings.grid_ms = ms;
	c->build_timed = 0;
}

static int render_passes(fr_context* ctx, int passes);

// every host wait of a context ends here: the stream is drained and the end-of-build status of the frames built since
// the last wait is read (resolve_frame).  FR_RETRIED: a frame that had only been built, not rendered, found the tables
// of its slot too small and has been rebuilt.
static int finish_pending_ex(Context* c)
{
	c->wait_epoch++;                     // (as stream_sync: whoever calls this may have waited)
	bool const rendered = c->render_pending;
	if (c->render_pending)
	{
		// (a lane records no completion event: it waits for its whole stream, copies behind the render included)
		if (c->blocking_sync || !c->stage_timing) { int const rc = stream_sync(c); if (rc) return rc; }
		else FM_CUDA(cudaEventSynchronize(c->ev_done));
		c->render_pending = false;
	}
	int retried = FR_OK;
	for (size_t k = 0; k < c->pending_frames.size(); k++)
	{
		int const idx = c->pending_frames[k];
		if (idx < 0 || (size_t)idx >= c->frames.size()) continue;
		int const rc = resolve_frame(c, &c->frames[idx], rendered);
		if (rc < 0) { c->pending_frames.clear(); c->build_timed = 0; return rc; }
		if (rc == FR_RETRIED) retried = FR_RETRIED;
	}
	c->pending_frames.clear();
	if (rendered)
	{
		read_build_timings(c);
		float ms = 0.0f;
		if (c->stage_timing && cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]) == cudaSuccess) c->timings.depth_ms = ms;
		if (c->stage_timing && cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]) == cudaSuccess) c->timings.march_ms = ms;
		if (c->march_timed && c->stage_timing)
		{
			if (cudaEventElapsedTime(&ms, c->ev[5], c->ev[10]) == cudaSuccess) c->timings.classify_ms = ms;
			if (cudaEventElapsedTime(&ms, c->ev[10], c->ev[11]) == cudaSuccess) c->timings.march_first_ms = ms;
			if (cudaEventElapsedTime(&ms, c->ev[11], c->ev[6]) == cudaSuccess) c->timings.march_long_ms = ms;
		}
	}
	else if (c->build_timed)       // a frame build without a render behind it: its times are read once they exist
	{
		if (cudaEventQuery(c->ev[3]) == cudaSuccess) read_build_timings(c);
	}
	return retried;
}

static int finish_pending(Context* c)
{
	int const rc = finish_pending_ex(c);
	return rc == FR_RETRIED ? FR_OK : rc;
}

// host copy of a frame's grid parameters for the entry points that need them
static int resolved(Context* c, Frame* f)
{
	int const rc = resolve_frame(c, f);
	return rc == FR_RETRIED ? FR_OK : rc;
}

}  // namespace fm

using namespace fm;


#define FR_CHECK_CTX(ctx)                                                    \
	do {                                                                     \
		if (!(ctx)) { fm::set_error("null context"); return FR_ERR_INVALID; } \
		FM_CUDA(cudaSetDevice((ctx)->device));                               \
	} while (0)

extern "C" {

int fr_abi_version(void) { return FR_ABI_VERSION; }

const char* fr_last_error(void) { return fm::g_error.c_str(); }

int fr_create(int device, int width, int height, fr_context** out)
{
	if (!out || width <= 0 || height <= 0) { set_error("fr_create: bad arguments"); return FR_ERR_INVALID; }
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
	{
		cudaGetLastError();
		set_error(std::string("fr_create: no CUDA device (") + cudaGetErrorString(e) +
				  "); fluidmarch has no CPU fallback");
		return FR_ERR_NO_DEVICE;
	}
	if (device < 0 || device >= count) { set_error("fr_create: device index out of range"); return FR_ERR_INVALID; }
	cudaDeviceProp prop;
	FM_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10)
	{
		set_error("fr_create: device is not sm_100 or newer; the kernels are built for sm_100a only");
		return FR_ERR_NO_DEVICE;
	}
	FM_CUDA(cudaSetDevice(device));
	fr_context* c = new (std::nothrow) fr_context();
	if (!c) { set_error("out of host memory"); return FR_ERR_INVALID; }
	c->device = device;
	c->sm_count = prop.multiProcessorCount;
	int rc = FR_OK;
	do
	{
		if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		if (cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		if (cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		if (cudaStreamCreateWithFlags(&c->stream_depth, cudaStreamNonBlocking) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		if (cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		if (cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		if (cudaEventCreateWithFlags(&c->ev_fork2, cudaEventDisableTiming) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		if (cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		if (const char* e = getenv("FLUIDMARCH_OVERLAP")) c->overlap_depth = e[0] != '0';
		if (cudaHostAlloc((void**)&c->h_sync_flag, 64, cudaHostAllocMapped) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		*c->h_sync_flag = 0u;
		if (cudaHostGetDevicePointer((void**)&c->d_sync_flag, (void*)c->h_sync_flag, 0) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		for (auto& ev : c->ev)
			if (cudaEventCreate(&ev) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		if (rc) break;
		if (cudaMalloc((void**)&c->d_counters, sizeof(DeviceCounters)) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		if (cudaMemset(c->d_counters, 0, sizeof(DeviceCounters)) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		if (cudaMallocHost((void**)&c->h_counters, sizeof(DeviceCounters)) != cudaSuccess) { rc = FR_ERR_CUDA; break; }
		rc = alloc_images(c, width, height);
	} while (0);
	if (rc != FR_OK)
	{
		if (rc == FR_ERR_CUDA && g_error.empty()) cuda_fail(cudaGetLastError(), "fr_create", __FILE__, __LINE__);
		fr_destroy(c);
		return rc;
	}
	// VisualizationSettings defaults (AdvancedRenderer.cpp:18-28) with the isotropic kernel
	c->settings.frame = 0;
	c->settings.max_steps = 128;
	c->settings.step_size = 0.009f;
	c->settings.iso_density = 1.0f;
	c->settings.enable_anisotropy = 0;
	c->settings.k_n = 0.5f; c->settings.k_r = 2.0f; c->settings.k_s = 2000.0f; c->settings.n_eps = 1;
	c->have_settings = true;
	if (const char* e = getenv("FR_DEPTH_REFINE")) c->depth_refine_bounds = e[0] == '1';   // tuning switch, same image either way
	*out = c;
	return FR_OK;
}

int fr_resize(fr_context* ctx, int width, int height)
{
	FR_CHECK_CTX(ctx);
	if (width <= 0 || height <= 0) { set_error("fr_resize: bad size"); return FR_ERR_INVALID; }
	{ int const src = stream_sync(ctx); if (src) return src; }
	ctx->render_pending = false;
	bool const external = ctx->d_rgba_target != ctx->d_rgba;
	if (ctx->ext_mem_pos) { cudaDestroyExternalMemory(ctx->ext_mem_pos); ctx->ext_mem_pos = nullptr; }       // imported images are
	if (ctx->ext_mem_nrm) { cudaDestroyExternalMemory(ctx->ext_mem_nrm); ctx->ext_mem_nrm = nullptr; }       // tied to the old size
	free_images(ctx);
	int rc = alloc_images(ctx, width, height);
	if (external) ctx->d_rgba_target = ctx->d_rgba;   // an external target is tied to the old size
	return rc;
}

void fr_destroy(fr_context* ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	for (auto& f : ctx->frames) free_frame(f);
	free_images(ctx);
	if (ctx->d_xyz) cudaFree(ctx->d_xyz);
	if (ctx->peer_rgba) cudaIpcCloseMemHandle(ctx->peer_rgba);
	if (ctx->d_raw) cudaFree(ctx->d_raw);
	if (ctx->d_bmp) cudaFree(ctx->d_bmp);
	if (ctx->h_bmp) cudaFreeHost(ctx->h_bmp);
	if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
	if (ctx->d_keys) cudaFree(ctx->d_keys);
	if (ctx->d_sort_tmp) cudaFree(ctx->d_sort_tmp);
	if (ctx->d_scan_tmp) cudaFree(ctx->d_scan_tmp);
	if (ctx->d_aabb_partial) cudaFree(ctx->d_aabb_partial);
	if (ctx->d_tile_bound) cudaFree(ctx->d_tile_bound);
	if (ctx->d_splat) cudaFree(ctx->d_splat);
	if (ctx->d_tiles) cudaFree(ctx->d_tiles);
	if (ctx->d_rayq) cudaFree(ctx->d_rayq);
	if (ctx->d_survivors) cudaFree(ctx->d_survivors);
	if (ctx->d_tmp_idx) cudaFree(ctx->d_tmp_idx);
	if (ctx->d_smooth) cudaFree(ctx->d_smooth);
	if (ctx->d_counters) cudaFree(ctx->d_counters);
	if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
	if (ctx->ext_wait) cudaDestroyExternalSemaphore(ctx->ext_wait);
	if (ctx->ext_signal) cudaDestroyExternalSemaphore(ctx->ext_signal);
	if (ctx->ext_mem) cudaDestroyExternalMemory(ctx->ext_mem);
	if (ctx->ext_mem_pos) cudaDestroyExternalMemory(ctx->ext_mem_pos);
	if (ctx->ext_mem_nrm) cudaDestroyExternalMemory(ctx->ext_mem_nrm);
	for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
	if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
	if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
	if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
	if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
	if (ctx->ev_fork2) cudaEventDestroy(ctx->ev_fork2);
	if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
	if (ctx->stream_depth) cudaStreamDestroy(ctx->stream_depth);
	if (ctx->h_sync_flag) cudaFreeHost((void*)ctx->h_sync_flag);
	if (ctx->stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
}

int fr_host_alloc(size_t bytes, void** out)
{
	if (!out) { set_error("fr_host_alloc: null out"); return FR_ERR_INVALID; }
	FM_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
	return FR_OK;
}

void fr_host_free(void* p)
{
	if (p) cudaFreeHost(p);
}

// ---- frames --------------------------------------------------------------------------------------

static int frame_slot(fr_context* ctx, int frame, Frame** out)
{
	if (frame < 0 || frame > (1 << 20)) { set_error("frame index out of range"); return FR_ERR_INVALID; }
	if ((size_t)frame >= ctx->frames.size()) ctx->frames.resize((size_t)frame + 1);
	*out = &ctx->frames[frame];
	return FR_OK;
}

int fr_upload_frame(fr_context* ctx, int frame, const float* xyz_host, size_t n, float h, float h_ext_mult)
{
	FR_CHECK_CTX(ctx);
	if (!xyz_host) { set_error("fr_upload_frame: null particle array"); return FR_ERR_INVALID; }
	if (n == 0) { set_error("fr_upload_frame: a frame needs at least one particle (Frame::ComputeAABB reads m_Particles[0])"); return FR_ERR_INVALID; }
	Frame* f;
	int rc = frame_slot(ctx, frame, &f);
	if (rc) return rc;
	if ((rc = finish_pending(ctx))) return rc;
	if ((rc = ensure_capacity(&ctx->d_xyz, &ctx->cap_xyz, n * 3))) return rc;
	FM_TIME(ctx, ctx->ev[0], ctx->stream);
	FM_CUDA(cudaMemcpyAsync(ctx->d_xyz, xyz_host, n * 12, cudaMemcpyHostToDevice, ctx->stream));
	FM_TIME(ctx, ctx->ev[1], ctx->stream);
	FM_CUDA(cudaEventRecord(ctx->ev_copy, ctx->stream));
	rc = build_frame(ctx, f, ctx->d_xyz, n, h, h_ext_mult);
	if (rc) return rc;
	ctx->pending_frames.push_back(frame);
	ctx->build_timed = 2;
	// The build stays queued on the stream (a build into the tables of an earlier frame never waited for the device at
	// all), so a render can be queued right behind it; xyz_host, however, is the caller's again when this returns
	FM_CUDA(cudaEventSynchronize(ctx->ev_copy));
	return FR_OK;
}

int fr_set_async_build(fr_context* ctx, int on)
{
	FR_CHECK_CTX(ctx);
	ctx->async_build = on != 0;
	return FR_OK;
}

int fr_set_stage_timing(fr_context* ctx, int on)
{
	FR_CHECK_CTX(ctx);
	int const rc = fm::finish_pending(ctx);
	if (rc) return rc;
	ctx->stage_timing = on != 0;
	if (!ctx->stage_timing) { ctx->build_timed = 0; ctx->march_timed = false; ctx->timings = fr_timings{}; }
	return FR_OK;
}

}  // extern "C"

// ---- a sequence lane's frame in two halves (fm_sequence.cu): everything up to the one host wait of the frame, and
// ---- the rest, which is launches and copies only and can therefore be captured into a CUDA graph
namespace fm
{

// A lane's frame in three steps.  (1) lane_frame_begin: whatever needs the host -- a particle file is read and staged;
// decides whether the rest can be queued without a wait.  (2) lane_frame_stage_a: upload, frame build, depth pre-pass --
// (the upload goes first, by itself) launches only when the build does not wait (every frame of a lane but its
// first), so the lane captures it into a CUDA graph: one stream operation.  Then lane_frame_resolve picks the grid parameters up from mapped memory (no stream
// operation at all).  (3) lane_frame_enqueue: march and copies out, the lane's second graph.
int lane_frame_begin(fr_context* ctx, const fr_seq_job& job, const char* bgeo_path, bool* stage_a_is_launches_only)
{
	FM_CUDA(cudaSetDevice(ctx->device));
	Frame* f;
	int rc = frame_slot(ctx, 0, &f);
	if (rc) return rc;
	if ((rc = finish_pending(ctx))) return rc;
	ctx->lane_d_xyz = nullptr;
	ctx->lane_n = (size_t)job.n;
	ctx->lane_h2d = false;
	if (bgeo_path)
	{
		if ((rc = stage_bgeo(ctx, bgeo_path, &ctx->lane_n))) return rc;
		ctx->lane_d_xyz = ctx->d_xyz;
	}
	else if (job.xyz_on_device) ctx->lane_d_xyz = job.xyz;
	else
	{
		if (!job.xyz || ctx->lane_n == 0) { set_error("sequence job without particles"); return FR_ERR_INVALID; }
		if ((rc = ensure_capacity(&ctx->d_xyz, &ctx->cap_xyz, ctx->lane_n * 3))) return rc;
		ctx->lane_d_xyz = ctx->d_xyz;
		ctx->lane_h2d = true;
	}
	ctx->build_timed = 0;
	if (!ctx->have_camera) { set_error("sequence: fr_seq_set_camera has not been called"); return FR_ERR_STATE; }
	int const passes = job.passes ? job.passes : FR_PASS_ALL;
	if ((passes & FR_PASS_MARCH) && !(passes & FR_PASS_DEPTH) && !ctx->have_depth)
	{
		set_error("sequence: march without a depth image");
		return FR_ERR_STATE;
	}
	*stage_a_is_launches_only = build_will_not_wait(ctx, f, true);
	return FR_OK;
}

// (the upload stays outside the lane's graph: the job's particles may sit in pageable memory)
int lane_frame_upload(fr_context* ctx, const fr_seq_job& job)
{
	if (ctx->lane_h2d) FM_CUDA(cudaMemcpyAsync(ctx->d_xyz, job.xyz, ctx->lane_n * 12, cudaMemcpyHostToDevice, ctx->stream));
	return FR_OK;
}

int lane_frame_stage_a(fr_context* ctx, const fr_seq_job& job)
{
	Frame* const f = &ctx->frames[0];
	int rc;
	if ((rc = build_frame(ctx, f, ctx->lane_d_xyz, ctx->lane_n, job.h, job.h_ext_mult, true))) return rc;
	ctx->pending_frames.push_back(0);
	return render_depth(ctx, job.passes ? job.passes : FR_PASS_ALL, false);
}

int lane_frame_resolve(fr_context* ctx, const fr_seq_job& job)
{
	return render_resolve(ctx, job.passes ? job.passes : FR_PASS_ALL);
}

// the rest of the frame -- march, copies out -- is launches only: a lane captures it into its CUDA graph
// (stream-ordered store of a frame's completion flag; the flag may live in another GPU's memory)
__global__ void k_signal(uint32_t* flag, uint32_t value)
{
	__threadfence_system();
	*reinterpret_cast<volatile uint32_t*>(flag) = value;
}

int lane_frame_enqueue(fr_context* ctx, const fr_seq_job& job)
{
	int rc;
	uchar4* const own_target = ctx->d_rgba_target;
	if (job.rgba_device) ctx->d_rgba_target = (uchar4*)job.rgba_device;         // this frame only
	rc = render_march(ctx, job.passes ? job.passes : FR_PASS_ALL);
	cudaStream_t const s = ctx->stream;
	if (rc == FR_OK && job.done_flag_device)
	{
		k_signal<<<1, 1, 0, s>>>(job.done_flag_device, job.done_value);
		ctx->kernel_launches += 1;
		if (cudaGetLastError() != cudaSuccess) rc = FR_ERR_CUDA;
	}
	if (rc) { ctx->d_rgba_target = own_target; return rc; }
	size_t const npix = (size_t)ctx->width * ctx->height;
	if (job.depth) FM_CUDA(cudaMemcpyAsync(job.depth, ctx->d_depth, npix * 4, cudaMemcpyDeviceToHost, s));
	if (job.positions) FM_CUDA(cudaMemcpyAsync(job.positions, ctx->d_pos, npix * 16, cudaMemcpyDeviceToHost, s));
	if (job.normals) FM_CUDA(cudaMemcpyAsync(job.normals, ctx->d_nrm, npix * 16, cudaMemcpyDeviceToHost, s));
	if (job.rgba) FM_CUDA(cudaMemcpyAsync(job.rgba, ctx->d_rgba_target, npix * 4, cudaMemcpyDeviceToHost, s));
	ctx->d_rgba_target = own_target;
	return FR_OK;
}

int lane_frame_wait(fr_context* ctx)
{
	FM_CUDA(cudaSetDevice(ctx->device));
	return finish_pending(ctx);
}

}  // namespace fm

extern "C" {

int fr_build_frame_device(fr_context* ctx, int frame, const float* xyz_device, size_t n, float h, float h_ext_mult)
{
	FR_CHECK_CTX(ctx);
	if (!xyz_device || n == 0) { set_error("fr_build_frame_device: empty particle array"); return FR_ERR_INVALID; }
	Frame* f;
	int rc = frame_slot(ctx, frame, &f);
	if (rc) return rc;
	if ((rc = finish_pending(ctx))) return rc;
	rc = build_frame(ctx, f, xyz_device, n, h, h_ext_mult);
	if (rc) return rc;
	ctx->pending_frames.push_back(frame);
	ctx->build_timed = 1;
	return FR_OK;                  // xyz_device must stay valid until the host next waits for this context
}

int fr_set_count_mode(fr_context* ctx, int mode)
{
	FR_CHECK_CTX(ctx);
	if (mode != FR_COUNT_CELL_EXACT && mode != FR_COUNT_CENTRE_BOX) { set_error("fr_set_count_mode: unknown mode"); return FR_ERR_INVALID; }
	ctx->count_mode = mode;
	return FR_OK;
}

int fr_get_frame_info(fr_context* ctx, int frame, fr_frame_info* out)
{
	FR_CHECK_CTX(ctx);
	if (!out) { set_error("fr_get_frame_info: null out"); return FR_ERR_INVALID; }
	Frame* f = get_frame(ctx, frame, true);
	if (!f) return FR_ERR_STATE;
	{ int const rrc = resolved(ctx, f); if (rrc) return rrc; }
	unsigned long long occ = 0;
	FM_CUDA(cudaMemcpyAsync(&occ, f->d_occupied, sizeof occ, cudaMemcpyDeviceToHost, ctx->stream));
	{ int const src = stream_sync(ctx); if (src) return src; }
	memset(out, 0, sizeof *out);
	out->num_particles = f->n;
	out->h = f->h;
	for (int a = 0; a < 3; a++)
	{
		out->min[a] = f->gp.mn[a];
		out->max[a] = f->gp.mx[a];
		out->grid_dims[a] = f->gp.gdim[a];
		out->search_min[a] = f->gp.kmin[a];
		out->search_dims[a] = f->gp.kdim[a];
	}
	out->occupied_cells = occ;
	return FR_OK;
}

int fr_release_frame(fr_context* ctx, int frame)
{
	FR_CHECK_CTX(ctx);
	Frame* f = get_frame(ctx, frame, false);
	if (!f) return FR_ERR_INVALID;
	{ int const src = stream_sync(ctx); if (src) return src; }
	free_frame(*f);
	return FR_OK;
}

int fr_download_frame(fr_context* ctx, int frame, float* sorted_xyzi, uint32_t* cell_start,
					  uint32_t* grid_counts, uint8_t* grid_flags)
{
	FR_CHECK_CTX(ctx);
	Frame* f = get_frame(ctx, frame, true);
	if (!f) return FR_ERR_STATE;
	{ int const rrc = resolved(ctx, f); if (rrc) return rrc; }
	cudaStream_t const s = ctx->stream;
	size_t const cells = (size_t)f->gp.kdim[0] * f->gp.kdim[1] * f->gp.kdim[2];
	size_t const gcells = (size_t)f->gp.gdim[0] * f->gp.gdim[1] * f->gp.gdim[2];
	if (sorted_xyzi) FM_CUDA(cudaMemcpyAsync(sorted_xyzi, f->d_sorted, f->n * 16, cudaMemcpyDeviceToHost, s));
	if (cell_start) FM_CUDA(cudaMemcpyAsync(cell_start, f->d_cell_start, (cells + 1) * 4, cudaMemcpyDeviceToHost, s));
	if (grid_counts) FM_CUDA(cudaMemcpyAsync(grid_counts, f->d_grid_counts, gcells * 4, cudaMemcpyDeviceToHost, s));
	std::vector<uint32_t> words;
	if (grid_flags)
	{
		words.resize((gcells + 31) / 32);
		FM_CUDA(cudaMemcpyAsync(words.data(), f->d_occ_bits, words.size() * 4, cudaMemcpyDeviceToHost, s));
	}
	{ int const src = stream_sync(ctx); if (src) return src; }
	if (grid_flags)
		for (size_t c = 0; c < gcells; c++) grid_flags[c] = (uint8_t)((words[c >> 5] >> (c & 31)) & 1u);
	return FR_OK;
}

// ---- per-render state ----------------------------------------------------------------------------

int fr_set_settings(fr_context* ctx, const fr_settings* s)
{
	FR_CHECK_CTX(ctx);
	if (!s) { set_error("fr_set_settings: null"); return FR_ERR_INVALID; }
	if (s->max_steps < 0 || s->bisection_steps < 0 || s->bisection_steps > 64)
	{
		set_error("fr_set_settings: max_steps / bisection_steps out of range");
		return FR_ERR_INVALID;
	}
	if (s->enable_anisotropy && s->n_eps < 0)
	{
		set_error("fr_set_settings: n_eps must not be negative");
		return FR_ERR_INVALID;
	}
	ctx->settings = *s;
	ctx->have_settings = true;
	return FR_OK;
}

int fr_set_camera(fr_context* ctx, const fr_camera* cam)
{
	FR_CHECK_CTX(ctx);
	if (!cam) { set_error("fr_set_camera: null"); return FR_ERR_INVALID; }
	ctx->camera = *cam;
	ctx->have_camera = true;
	return FR_OK;
}

int fr_set_depth(fr_context* ctx, const float* depth_host)
{
	FR_CHECK_CTX(ctx);
	if (!depth_host) { set_error("fr_set_depth: null"); return FR_ERR_INVALID; }
	int rc = finish_pending(ctx);
	if (rc) return rc;
	size_t const npix = (size_t)ctx->width * ctx->height;
	FM_CUDA(cudaMemcpyAsync(ctx->d_depth, depth_host, npix * 4, cudaMemcpyHostToDevice, ctx->stream));
	ctx->have_depth = true;
	return FR_OK;
}

int fr_set_tile_partition(fr_context* ctx, int rank, int world, int tile_w, int tile_h)
{
	FR_CHECK_CTX(ctx);
	if (world < 1 || rank < 0 || rank >= world || tile_w < 32 || tile_h < 8 || tile_w % 32 || tile_h % 8)
	{
		set_error("fr_set_tile_partition: need 0 <= rank < world, tile_w a multiple of 32, tile_h a multiple of 8");
		return FR_ERR_INVALID;
	}
	ctx->part_rank = rank; ctx->part_world = world; ctx->part_tw = tile_w; ctx->part_th = tile_h;
	return FR_OK;
}

int fr_set_region_partition(fr_context* ctx, int x0, int y0, int x1, int y1)
{
	FR_CHECK_CTX(ctx);
	if (x1 == 0 && y1 == 0 && x0 == 0 && y0 == 0) { ctx->region[0] = ctx->region[1] = ctx->region[2] = ctx->region[3] = 0; return FR_OK; }
	bool const ok = x0 >= 0 && y0 >= 0 && x1 > x0 && y1 > y0 && x1 <= ctx->width && y1 <= ctx->height && x0 % 64 == 0 && y0 % 64 == 0 &&
		(x1 % 64 == 0 || x1 == ctx->width) && (y1 % 64 == 0 || y1 == ctx->height);
	if (!ok)
	{
		set_error("fr_set_region_partition: need 0 <= x0 < x1 <= W, 0 <= y0 < y1 <= H, bounds multiples of 64 (or the image edge)");
		return FR_ERR_INVALID;
	}
	ctx->region[0] = x0; ctx->region[1] = y0; ctx->region[2] = x1; ctx->region[3] = y1;
	return FR_OK;
}

// ---- render ----------------------------------------------------------------------------------------

int fr_render_async(fr_context* ctx, int passes)
{
	FR_CHECK_CTX(ctx);
	if (!(passes & FR_PASS_ALL)) { set_error("fr_render_async: no pass selected"); return FR_ERR_INVALID; }
	if (!ctx->have_camera) { set_error("fr_render_async: fr_set_camera has not been called"); return FR_ERR_STATE; }
	Frame* f = get_frame(ctx, ctx->settings.frame, true);
	if (!f) return FR_ERR_STATE;
	if ((passes & FR_PASS_MARCH) && !(passes & FR_PASS_DEPTH) && !ctx->have_depth)
	{
		set_error("fr_render_async: march without a depth image (call fr_set_depth or include FR_PASS_DEPTH)");
		return FR_ERR_STATE;
	}
	// a render still in flight is finished first (its timings are read); a frame build that is merely queued is not
	// waited for: the render goes right behind it on the stream
	if (ctx->render_pending) { int const rc = finish_pending(ctx); if (rc) return rc; }
	return render_passes(ctx, passes);
}

}  // extern "C"

namespace fm
{

// the part of a render that does not need the frame's grid parameters on the host: the depth pre-pass
int render_depth(fr_context* ctx, int passes, bool again)
{
	Frame* f = get_frame(ctx, ctx->settings.frame, true);
	if (!f) return FR_ERR_STATE;
	int rc;
	cudaStream_t const s = ctx->stream;
	if (ctx->ext_wait && !again)
	{
		cudaExternalSemaphoreWaitParams wp;
		memset(&wp, 0, sizeof wp);
		FM_CUDA(cudaWaitExternalSemaphoresAsync(&ctx->ext_wait, &wp, 1, s));
	}
	FM_TIME(ctx, ctx->ev[4], s);
	ctx->zero_counters_in_depth = (passes & FR_PASS_DEPTH) && (passes & (FR_PASS_MARCH | FR_PASS_SHADE));
	if (passes & FR_PASS_DEPTH)
	{
		// Behind a frame build that is still queued the pre-pass runs BESIDE it, on the second stream, reading the raw
		// particle array the build reads: the image is a minimum over all fragments, so the particle order does not
		// matter, and neither pass fills the GPU by itself (C2: build 0.063 ms, pre-pass 0.098 ms, one after the other
		// on one stream).  Not with per-stage events (their times would overlap), a region partition (the sorted array
		// holds only the region's particles) or an external semaphore in front.
		bool const region = ctx->region[2] > ctx->region[0];
		bool const beside = ctx->overlap_depth && !ctx->stage_timing && !again && !ctx->ext_wait && !region && !f->filtered && f->src_xyz &&
			ctx->fork_serial == f->build_serial && f->src_epoch == ctx->wait_epoch;
		if (beside)
		{
			FM_CUDA(cudaStreamWaitEvent(ctx->stream_depth, ctx->ev_fork, 0));
			if ((rc = launch_depth_prepass(ctx, *f, ctx->stream_depth, f->src_xyz))) return rc;
			FM_CUDA(cudaEventRecord(ctx->ev_join, ctx->stream_depth));
			FM_CUDA(cudaStreamWaitEvent(s, ctx->ev_join, 0));
		}
		else if ((rc = launch_depth_prepass(ctx, *f, s, nullptr))) return rc;
		ctx->have_depth = true;
	}
	FM_TIME(ctx, ctx->ev[5], s);
	return FR_OK;
}

// the part that does (the march kernels take the frame's geometry as kernel parameters): resolve_early comes first
int render_march(fr_context* ctx, int passes)
{
	Frame* f = get_frame(ctx, ctx->settings.frame, true);
	if (!f) return FR_ERR_STATE;
	int rc;
	cudaStream_t const s = ctx->stream;
	// Frame::m_SearchExt (the reference builds it in Frame::Frame; here on the frame's first anisotropic render)
	if ((passes & FR_PASS_MARCH) && ctx->settings.enable_anisotropy && (rc = build_frame_ext(ctx, f))) return rc;
	ctx->march_timed = (passes & (FR_PASS_MARCH | FR_PASS_SHADE)) != 0;
	if (ctx->march_timed)
		if ((rc = launch_march(ctx, *f, (passes & FR_PASS_MARCH) != 0, (passes & FR_PASS_SHADE) != 0))) return rc;
	FM_TIME(ctx, ctx->ev[6], s);
	if (ctx->ext_signal)
	{
		cudaExternalSemaphoreSignalParams sp;
		memset(&sp, 0, sizeof sp);
		FM_CUDA(cudaSignalExternalSemaphoresAsync(&ctx->ext_signal, &sp, 1, s));
	}
	if (ctx->stage_timing) FM_CUDA(cudaEventRecord(ctx->ev_done, s));
	ctx->render_pending = true;
	ctx->last_passes = passes;
	return FR_OK;
}

// depth pre-pass (queued behind the build without waiting), then the frame's grid parameters from the side stream --
// long there by now: the rest of the build and the pre-pass are still running --, then the march
int render_resolve(fr_context* ctx, int passes)
{
	Frame* f = get_frame(ctx, ctx->settings.frame, true);
	if (!f) return FR_ERR_STATE;
	int rc = resolve_early(ctx, f);
	if (rc < 0) return rc;
	if (rc == FR_RETRIED)
	{
		ctx->pending_frames.push_back(ctx->settings.frame);
		if ((rc = render_depth(ctx, passes, true))) return rc;      // the pre-pass ran on the unusable first attempt
	}
	return FR_OK;
}

static int render_passes(fr_context* ctx, int passes)
{
	int rc = render_depth(ctx, passes, false);
	if (rc) return rc;
	if ((rc = render_resolve(ctx, passes))) return rc;
	return render_march(ctx, passes);
}

}  // namespace fm

extern "C" {

int fr_is_done(fr_context* ctx)
{
	FR_CHECK_CTX(ctx);
	if (!ctx->render_pending) return 1;
	cudaError_t const e = ctx->stage_timing ? cudaEventQuery(ctx->ev_done) : cudaStreamQuery(ctx->stream);
	if (e == cudaSuccess) return 1;
	if (e == cudaErrorNotReady) return 0;
	return cuda_fail(e, "cudaEventQuery", __FILE__, __LINE__);
}

int fr_wait(fr_context* ctx)
{
	FR_CHECK_CTX(ctx);
	return finish_pending(ctx);
}

// ---- results -----------------------------------------------------------------------------------------

int fr_download(fr_context* ctx, float* depth, float* positions, float* normals, uint8_t* rgba)
{
	FR_CHECK_CTX(ctx);
	int rc = finish_pending(ctx);
	if (rc) return rc;
	cudaStream_t const s = ctx->stream;
	size_t const npix = (size_t)ctx->width * ctx->height;
	FM_TIME(ctx, ctx->ev[7], s);
	if (depth) FM_CUDA(cudaMemcpyAsync(depth, ctx->d_depth, npix * 4, cudaMemcpyDeviceToHost, s));
	if (positions) FM_CUDA(cudaMemcpyAsync(positions, ctx->d_pos, npix * 16, cudaMemcpyDeviceToHost, s));
	if (normals) FM_CUDA(cudaMemcpyAsync(normals, ctx->d_nrm, npix * 16, cudaMemcpyDeviceToHost, s));
	if (rgba) FM_CUDA(cudaMemcpyAsync(rgba, ctx->d_rgba_target, npix * 4, cudaMemcpyDeviceToHost, s));
	FM_TIME(ctx, ctx->ev[8], s);
	{ int const src = stream_sync(ctx); if (src) return src; }
	float ms = 0.0f;
	if (ctx->stage_timing && cudaEventElapsedTime(&ms, ctx->ev[7], ctx->ev[8]) == cudaSuccess) ctx->timings.download_ms = ms;
	return FR_OK;
}

int fr_device_images(fr_context* ctx, void** depth, void** positions, void** normals, void** rgba)
{
	FR_CHECK_CTX(ctx);
	if (depth) *depth = ctx->d_depth;
	if (positions) *positions = ctx->d_pos;
	if (normals) *normals = ctx->d_nrm;
	if (rgba) *rgba = ctx->d_rgba_target;
	return FR_OK;
}

int fr_set_color_target(fr_context* ctx, void* rgba_device)
{
	FR_CHECK_CTX(ctx);
	int rc = finish_pending(ctx);
	if (rc) return rc;
	ctx->d_rgba_target = rgba_device ? (uchar4*)rgba_device : ctx->d_rgba;
	return FR_OK;
}

// ---- tile-parallel over peer memory ---------------------------------------------------------------------------
// The presenting GPU exports its colour image; every other rank opens it and renders its own screen tiles straight
// into it: the shading epilogue's stores travel over NVLink / NVSwitch, there is no separate gather.
int fr_ipc_export_color(fr_context* ctx, fr_ipc_handle* out)
{
	FR_CHECK_CTX(ctx);
	if (!out) { set_error("fr_ipc_export_color: null out"); return FR_ERR_INVALID; }
	static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(fr_ipc_handle), "fr_ipc_handle too small");
	cudaIpcMemHandle_t h;
	FM_CUDA(cudaIpcGetMemHandle(&h, ctx->d_rgba));          // the internal image is its own cudaMalloc allocation
	memset(out, 0, sizeof *out);
	memcpy(out->bytes, &h, sizeof h);
	return FR_OK;
}

int fr_ipc_open_color_target(fr_context* ctx, const fr_ipc_handle* handle)
{
	FR_CHECK_CTX(ctx);
	if (!handle) { set_error("fr_ipc_open_color_target: null handle"); return FR_ERR_INVALID; }
	int rc = finish_pending(ctx);
	if (rc) return rc;
	if (ctx->peer_rgba) { cudaIpcCloseMemHandle(ctx->peer_rgba); ctx->peer_rgba = nullptr; ctx->d_rgba_target = ctx->d_rgba; }
	cudaIpcMemHandle_t h;
	memcpy(&h, handle->bytes, sizeof h);
	void* p = nullptr;
	FM_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	ctx->peer_rgba = p;
	ctx->d_rgba_target = (uchar4*)p;
	return FR_OK;
}

int fr_ipc_close_color_target(fr_context* ctx)
{
	FR_CHECK_CTX(ctx);
	int rc = finish_pending(ctx);
	if (rc) return rc;
	if (ctx->peer_rgba)
	{
		{ int const src = stream_sync(ctx); if (src) return src; }
		FM_CUDA(cudaIpcCloseMemHandle(ctx->peer_rgba));
		ctx->peer_rgba = nullptr;
		ctx->d_rgba_target = ctx->d_rgba;
	}
	return FR_OK;
}

int fr_device_alloc(int device, size_t bytes, void** out)
{
	if (!out || bytes == 0) { set_error("fr_device_alloc: bad arguments"); return FR_ERR_INVALID; }
	FM_CUDA(cudaSetDevice(device));
	FM_CUDA(cudaMalloc(out, bytes));
	FM_CUDA(cudaMemset(*out, 0, bytes));
	return FR_OK;
}

int fr_device_free(int device, void* p)
{
	FM_CUDA(cudaSetDevice(device));
	if (p) FM_CUDA(cudaFree(p));
	return FR_OK;
}

int fr_ipc_export_buffer(int device, void* device_ptr, fr_ipc_handle* out)
{
	if (!device_ptr || !out) { set_error("fr_ipc_export_buffer: null argument"); return FR_ERR_INVALID; }
	FM_CUDA(cudaSetDevice(device));
	cudaIpcMemHandle_t h;
	FM_CUDA(cudaIpcGetMemHandle(&h, device_ptr));
	memset(out, 0, sizeof *out);
	memcpy(out->bytes, &h, sizeof h);
	return FR_OK;
}

int fr_ipc_open_buffer(int device, const fr_ipc_handle* handle, void** out)
{
	if (!handle || !out) { set_error("fr_ipc_open_buffer: null argument"); return FR_ERR_INVALID; }
	FM_CUDA(cudaSetDevice(device));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle->bytes, sizeof h);
	FM_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
	return FR_OK;
}

int fr_ipc_close_buffer(int device, void* p)
{
	FM_CUDA(cudaSetDevice(device));
	if (p) FM_CUDA(cudaIpcCloseMemHandle(p));
	return FR_OK;
}

int fr_get_counters(fr_context* ctx, fr_counters* out)
{
	FR_CHECK_CTX(ctx);
	if (!out) { set_error("fr_get_counters: null out"); return FR_ERR_INVALID; }
	int rc = finish_pending(ctx);
	if (rc) return rc;
	FM_CUDA(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, ctx->stream));
	{ int const src = stream_sync(ctx); if (src) return src; }
	const DeviceCounters& d = *ctx->h_counters;
	out->pixels = (uint64_t)ctx->width * (uint64_t)ctx->height;
	out->covered_rays = d.covered_rays;
	out->hit_rays = d.hit_rays;
	out->ray_steps = d.ray_steps;
	out->skip_iterations = d.skip_iterations;
	out->candidates = d.candidates;
	out->neighbours = d.neighbours;
	out->early_exits = d.early_exits;
	out->neighbour_overflow = d.neighbour_overflow;
	out->kernel_launches = ctx->kernel_launches;
	out->first_candidates = d.first_candidates;
	out->queued_rays = d.ctl[2];          // slots of the ray queue that were filled
	out->first_examined = d.first_examined;
	out->first_fallbacks = d.first_fallbacks;
	// (profiling builds, tools/build_variant.sh x -DFM_LONG_PROFILE: first_examined = long_ray's ray with the longest walk
	// (cycles << 32 | samples << 22 | skips << 10 | general-path skips), first_fallbacks = its slowest ray (cycles << 32 |
	// cycles evaluating))
	return FR_OK;
}

int fr_get_timings(fr_context* ctx, fr_timings* out)
{
	FR_CHECK_CTX(ctx);
	if (!out) { set_error("fr_get_timings: null out"); return FR_ERR_INVALID; }
	int rc = finish_pending(ctx);
	if (rc) return rc;
	*out = ctx->timings;
	return FR_OK;
}

int fr_get_stream(fr_context* ctx, void** stream)
{
	FR_CHECK_CTX(ctx);
	if (!stream) { set_error("fr_get_stream: null out"); return FR_ERR_INVALID; }
	*stream = (void*)ctx->stream;
	return FR_OK;
}

// ---- point queries -------------------------------------------------------------------------------------

int fr_query_neighbors(fr_context* ctx, int frame, const float* points_host, size_t m,
					   uint32_t* counts, uint32_t* ids, size_t cap)
{
	FR_CHECK_CTX(ctx);
	if ((m && (!points_host || !counts))) { set_error("fr_query_neighbors: null array"); return FR_ERR_INVALID; }
	Frame* f = get_frame(ctx, frame, true);
	if (!f) return FR_ERR_STATE;
	{ int const rrc = finish_pending(ctx); if (rrc) return rrc; }
	{ int const rrc = resolved(ctx, f); if (rrc) return rrc; }
	return query_neighbors(ctx, *f, points_host, m, counts, ids, cap, false);
}

int fr_query_neighbors_ext(fr_context* ctx, int frame, const float* points_host, size_t m,
						   uint32_t* counts, uint32_t* ids, size_t cap)
{
	FR_CHECK_CTX(ctx);
	if ((m && (!points_host || !counts))) { set_error("fr_query_neighbors_ext: null array"); return FR_ERR_INVALID; }
	Frame* f = get_frame(ctx, frame, true);
	if (!f) return FR_ERR_STATE;
	int rc = finish_pending(ctx);
	if (rc) return rc;
	if ((rc = resolved(ctx, f))) return rc;
	if ((rc = build_frame_ext(ctx, f))) return rc;
	return query_neighbors(ctx, *f, points_host, m, counts, ids, cap, true);
}

int fr_query_anisotropic(fr_context* ctx, int frame, const float* points_host, size_t m, float* density, float* grad, float* g9)
{
	FR_CHECK_CTX(ctx);
	if ((m && (!points_host || !density))) { set_error("fr_query_anisotropic: null array"); return FR_ERR_INVALID; }
	Frame* f = get_frame(ctx, frame, true);
	if (!f) return FR_ERR_STATE;
	int rc = finish_pending(ctx);
	if (rc) return rc;
	if ((rc = resolved(ctx, f))) return rc;
	if ((rc = build_frame_ext(ctx, f))) return rc;
	return query_aniso(ctx, *f, ctx->settings, points_host, m, density, grad, g9);
}

int fr_download_frame_ext(fr_context* ctx, int frame, float* sorted_xyzi, uint32_t* cell_start, int32_t search_min[3],
						  int32_t search_dims[3])
{
	FR_CHECK_CTX(ctx);
	Frame* f = get_frame(ctx, frame, true);
	if (!f) return FR_ERR_STATE;
	int rc = finish_pending(ctx);
	if (rc) return rc;
	if ((rc = resolved(ctx, f))) return rc;
	if ((rc = build_frame_ext(ctx, f))) return rc;
	cudaStream_t const s = ctx->stream;
	size_t const cells = (size_t)f->kdim_ext[0] * f->kdim_ext[1] * f->kdim_ext[2];
	if (sorted_xyzi) FM_CUDA(cudaMemcpyAsync(sorted_xyzi, f->d_sorted_ext, f->n * 16, cudaMemcpyDeviceToHost, s));
	if (cell_start) FM_CUDA(cudaMemcpyAsync(cell_start, f->d_cell_start_ext, (cells + 1) * 4, cudaMemcpyDeviceToHost, s));
	{ int const src = stream_sync(ctx); if (src) return src; }
	for (int a = 0; a < 3; a++)
	{
		if (search_min) search_min[a] = f->kmin_ext[a];
		if (search_dims) search_dims[a] = f->kdim_ext[a];
	}
	return FR_OK;
}

int fr_query_density(fr_context* ctx, int frame, const float* points_host, size_t m, float* density, float* grad)
{
	FR_CHECK_CTX(ctx);
	if ((m && (!points_host || !density))) { set_error("fr_query_density: null array"); return FR_ERR_INVALID; }
	Frame* f = get_frame(ctx, frame, true);
	if (!f) return FR_ERR_STATE;
	{ int const rrc = finish_pending(ctx); if (rrc) return rrc; }
	{ int const rrc = resolved(ctx, f); if (rrc) return rrc; }
	return query_density(ctx, *f, points_host, m, density, grad);
}

int fr_measure_l2_bandwidth(fr_context* ctx, size_t bytes, uint32_t reps, float* gbs)
{
	FR_CHECK_CTX(ctx);
	if (!gbs || reps == 0) { set_error("fr_measure_l2_bandwidth: bad arguments"); return FR_ERR_INVALID; }
	int rc = finish_pending(ctx);
	if (rc) return rc;
	return measure_l2_bandwidth(ctx, bytes, reps, gbs);
}

int fr_selftest_division(fr_context* ctx, uint64_t n, uint64_t seed, uint64_t* mismatches)
{
	FR_CHECK_CTX(ctx);
	if (!mismatches) { set_error("fr_selftest_division: null out"); return FR_ERR_INVALID; }
	return selftest_division(ctx, n, seed, mismatches);
}

// ---- CUDA-Vulkan hand-off ------------------------------------------------------------------------------

int fr_import_vk_memory_fd(fr_context* ctx, int fd, size_t allocation_bytes, size_t offset)
{
	FR_CHECK_CTX(ctx);
	size_t const need = (size_t)ctx->width * ctx->height * 4;
	if (fd < 0 || offset + need > allocation_bytes)
	{
		set_error("fr_import_vk_memory_fd: bad fd or the allocation is smaller than W*H*4 bytes at offset");
		return FR_ERR_INVALID;
	}
	int rc = finish_pending(ctx);
	if (rc) return rc;
	if (ctx->ext_mem) { cudaDestroyExternalMemory(ctx->ext_mem); ctx->ext_mem = nullptr; ctx->d_rgba_target = ctx->d_rgba; }
	cudaExternalMemoryHandleDesc hd;
	memset(&hd, 0, sizeof hd);
	hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
	hd.handle.fd = fd;                       // ownership of the fd passes to CUDA on success
	hd.size = allocation_bytes;
	FM_CUDA(cudaImportExternalMemory(&ctx->ext_mem, &hd));
	cudaExternalMemoryBufferDesc bd;
	memset(&bd, 0, sizeof bd);
	bd.offset = offset;
	bd.size = need;
	void* ptr = nullptr;
	FM_CUDA(cudaExternalMemoryGetMappedBuffer(&ptr, ctx->ext_mem, &bd));
	ctx->d_rgba_target = (uchar4*)ptr;
	return FR_OK;
}

static int import_fd_buffer(int fd, size_t allocation_bytes, size_t need, cudaExternalMemory_t* mem, void** ptr)
{
	cudaExternalMemoryHandleDesc hd;
	memset(&hd, 0, sizeof hd);
	hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
	hd.handle.fd = fd;                       // ownership of the fd passes to CUDA on success
	hd.size = allocation_bytes;
	FM_CUDA(cudaImportExternalMemory(mem, &hd));
	cudaExternalMemoryBufferDesc bd;
	memset(&bd, 0, sizeof bd);
	bd.offset = 0;
	bd.size = need;
	FM_CUDA(cudaExternalMemoryGetMappedBuffer(ptr, *mem, &bd));
	return FR_OK;
}

int fr_import_vk_images_fd(fr_context* ctx, int positions_fd, int normals_fd, size_t allocation_bytes)
{
	FR_CHECK_CTX(ctx);
	size_t const need = (size_t)ctx->width * ctx->height * 16;
	if (positions_fd < 0 || normals_fd < 0 || need > allocation_bytes)
	{
		set_error("fr_import_vk_images_fd: bad fd or the allocations are smaller than W*H*16 bytes");
		return FR_ERR_INVALID;
	}
	int rc = finish_pending(ctx);
	if (rc) return rc;
	if (ctx->ext_mem_pos) { cudaDestroyExternalMemory(ctx->ext_mem_pos); ctx->ext_mem_pos = nullptr; ctx->d_pos = ctx->d_pos_own; }
	if (ctx->ext_mem_nrm) { cudaDestroyExternalMemory(ctx->ext_mem_nrm); ctx->ext_mem_nrm = nullptr; ctx->d_nrm = ctx->d_nrm_own; }
	void* p = nullptr;
	if ((rc = import_fd_buffer(positions_fd, allocation_bytes, need, &ctx->ext_mem_pos, &p))) return rc;
	ctx->d_pos = (float4*)p;
	if ((rc = import_fd_buffer(normals_fd, allocation_bytes, need, &ctx->ext_mem_nrm, &p))) return rc;
	ctx->d_nrm = (float4*)p;
	return FR_OK;
}

int fr_import_vk_semaphores_fd(fr_context* ctx, int wait_fd, int signal_fd)
{
	FR_CHECK_CTX(ctx);
	int rc = finish_pending(ctx);
	if (rc) return rc;
	if (ctx->ext_wait) { cudaDestroyExternalSemaphore(ctx->ext_wait); ctx->ext_wait = nullptr; }
	if (ctx->ext_signal) { cudaDestroyExternalSemaphore(ctx->ext_signal); ctx->ext_signal = nullptr; }
	cudaExternalSemaphoreHandleDesc sd;
	if (wait_fd >= 0)
	{
		memset(&sd, 0, sizeof sd);
		sd.type = cudaExternalSemaphoreHandleTypeOpaqueFd;
		sd.handle.fd = wait_fd;
		FM_CUDA(cudaImportExternalSemaphore(&ctx->ext_wait, &sd));
	}
	if (signal_fd >= 0)
	{
		memset(&sd, 0, sizeof sd);
		sd.type = cudaExternalSemaphoreHandleTypeOpaqueFd;
		sd.handle.fd = signal_fd;
		FM_CUDA(cudaImportExternalSemaphore(&ctx->ext_signal, &sd));
	}
	return FR_OK;
}

}  // extern "C"
