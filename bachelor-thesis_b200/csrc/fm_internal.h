// fm_internal.h -- host-side structures shared by the fluidmarch translation units.
#pragma once

#include <cuda_runtime.h>
#include <utility>
#include <stdint.h>
#include <stddef.h>

#include <string>
#include <vector>

#include "../../include/fluidmarch.h"
#include "fm_common.cuh"

namespace fm
{

// written by the device (k_aabb_params) and read by every later kernel of the frame build from device memory, so that
// the build needs no host round trip; a copy comes back to the host with the frame's results (Frame::h_gp)
struct GridParams
{
	float mn[3], mx[3];          // m_Min, m_Max (Dataset.cpp:78-92)
	int32_t gdim[3];             // density grid W, H, D (Dataset.cpp:102-104)
	float cell_width;
	float inv_cell_width;
	int32_t kmin[3], kdim[3];    // neighbour-search cell range
	float search_inv;
	uint32_t raw_min[3], raw_max[3];   // order-preserving encodings of the particle extrema (kept for build_frame_ext)
	uint32_t cells, gcells;      // kdim / gdim products; 0 when status != 0
	uint32_t status;             // FM_GRID_* bits; != 0: the frame is unusable and every kernel of the build returns at once
	uint32_t max_cell;           // largest number of particles in one search cell (k_scan_flags)
	uint32_t n_sorted;           // particles in the sorted array: all of them, or those the region filter kept (k_scan_flags)
	uint32_t seq;                // number of the build that wrote this (the host polls it in the mapped copy, resolve_early)
};

// Region partition (tile-parallel multi-GPU, fr_set_region_partition): this context renders the pixel rectangle
// [x0, x1) x [y0, y1) only and builds its search structures from the particles that can influence those pixels: the ones
// within `margin` of the rectangle's view frustum (4 side planes through the camera position, unit normals pointing
// outward).  A sample of a ray of the region reads the 27 search cells around it, i.e. particles up to 2 h away per axis
// (2 sqrt(3) h in distance); the depth pre-pass needs the discs that overlap the rectangle (radius h).  With a margin of
// 3.6 h every cell such a sample can touch holds exactly the particles of the unpartitioned build, in the same order.
struct RegionFilter
{
	int on;
	float margin;
	float cam[3];
	float plane[4][3];
};

enum
{
	FM_GRID_NONFINITE = 1,       // a particle coordinate is NaN or infinite
	FM_GRID_DEGENERATE = 2,      // empty / non-finite bounds, or extent / h beyond 2^31 cells
	FM_GRID_OVERFLOW = 4         // the tables sized for an earlier frame are too small: rebuild with a host round trip
};
constexpr uint32_t kMaxCellParticles = 2048;     // above: the cell is sorted by one CTA instead of ranked quadratically (k_cell_order)

// table capacities the device checks the grid parameters against (all 0xffffffff: no check, the host sizes the tables
// after reading the parameters back)
struct BuildCaps
{
	uint32_t cells, gcells, occ_words, scan_tiles;
};

struct Frame
{
	bool valid = false;
	size_t n = 0;
	float h = 0.0f, h_ext = 0.0f;
	GridParams gp{};                  // host copy: valid once gp_pending is false (resolve_frame)
	GridParams* d_gp = nullptr;       // device: written by k_aabb_params, read by the build kernels
	GridParams* h_gp = nullptr;       // pinned + mapped, 2 entries: [0] written by k_aabb_params itself (zero-copy), [1] copied at the end of the build
	GridParams* h_gp_dev = nullptr;   // device address of h_gp[0]
	uint32_t gp_seq = 0;              // GridParams::seq of the queued build
	bool gp_early_pending = false;    // k_aabb_params of this build is queued: its parameters are on their way (resolve_early)
	bool gp_pending = false;          // the build is queued: its end-of-build status has not been read yet (resolve_frame)
	bool gp_host_valid = false;       // `gp` holds THIS build's parameters (after resolve_early; at once after a build with a host wait)
	uint64_t build_serial = 0;        // process-wide number of this build
	const float* src_xyz = nullptr;   // device particles of the queued build (must stay valid until the host next waits)
	uint64_t src_epoch = 0;           // Context::wait_epoch when the build was queued
	float src_mult = 0.0f;
	bool filtered = false;            // built under a region partition: holds only the particles its pixel rectangle needs
	uint64_t occupied = 0;
	// device buffers (capacities in elements)
	float4* d_sorted = nullptr;       size_t cap_sorted = 0;
	uint32_t* d_cell_start = nullptr; size_t cap_cells = 0;     // kdim product + 1
	uint32_t* d_grid_counts = nullptr; size_t cap_grid = 0;     // gdim product
	uint32_t* d_occ_bits = nullptr;   size_t cap_occ_words = 0;
	unsigned long long* d_occupied = nullptr;                   // number of flagged nodes
	// Frame::m_SearchExt (r = h_ext), built on demand by build_frame_ext
	bool ext_valid = false;
	int32_t kmin_ext[3] = {}, kdim_ext[3] = {};
	float search_inv_ext = 0.0f;
	float4* d_sorted_ext = nullptr;       size_t cap_sorted_ext = 0;
	uint32_t* d_cell_start_ext = nullptr; size_t cap_cells_ext = 0;
};

struct DeviceCounters
{
	unsigned long long covered_rays, hit_rays, ray_steps, skip_iterations, candidates, neighbours,
		early_exits, neighbour_overflow;
	unsigned long long first_candidates;   // the share of `candidates` examined by k_march_first
	unsigned long long first_examined;     // candidates k_march_first ran the distance test on (its staged walk culls cells)
	unsigned long long first_fallbacks;    // first samples walked out of global memory (tile did not fit the stage)
	uint32_t ctl[8];                       // work-list control words of the march (RayQueues::ctl), zeroed with the counters
};

struct Context
{
	int device = 0;
	int width = 0, height = 0;
	int sm_count = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev_done = nullptr;
	cudaEvent_t ev_copy = nullptr;     // behind the host -> device copy of fr_upload_frame
	// The depth pre-pass needs the particles, not the grid: with stage timing off (fr_set_stage_timing(ctx, 0): the
	// RayMarcher shim, latency measurements) a render queued behind a frame build runs it on a second stream, from the
	// raw particle array, while the build's kernels run on the first (render_depth).  ev_fork is recorded in front of
	// a build's kernels, ev_join behind the pre-pass.  wait_epoch counts the host waits of the context: the raw array
	// of a queued build is only guaranteed until the next one (fr_build_frame_device's contract).
	cudaStream_t stream_depth = nullptr;
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;      // the same around the uncovered pixels' outputs beside k_march_long (launch_march)
	uint64_t fork_serial = 0;          // Frame::build_serial of the build ev_fork belongs to
	uint64_t wait_epoch = 0;
	bool overlap_depth = true;         // FLUIDMARCH_OVERLAP=0 switches it off; sequence lanes: off (other lanes' frames fill the GPU)
	// Host waits.  By default a thread waits inside the driver (cudaStreamSynchronize spins: lowest latency, best
	// throughput while every waiting thread has a core of its own -- 0.296 ms per C2 frame at 6 lanes).  When the host
	// is oversubscribed (8 ranks x 6 lanes on a 32-core box: 0.47 ms) the lanes of a sequence can instead append a
	// stream memory operation (cuStreamWriteValue32) that writes a sequence number into mapped pinned memory and
	// poll that word, yielding the core between polls (0.42 ms there; 0.33 ms on an idle host), fr_seq_set_yielding.
	// Also tried (r01u): sleeping on a blocking-sync event (0.43 ms on an idle host: wake-up latency), polling
	// cudaEventQuery (0.33), a signalling kernel (waits for an SM behind the other lanes' persistent kernels).
	bool blocking_sync = false;
	volatile uint32_t* h_sync_flag = nullptr;
	uint32_t* d_sync_flag = nullptr;
	uint32_t sync_seq = 0;
	cudaEvent_t ev[16] = {};
	bool render_pending = false;
	// between build_frame_begin / _finish
	struct { Frame* f = nullptr; const float* d_xyz = nullptr; size_t n = 0; float h = 0.0f, h_ext_mult = 0.0f;
			 bool async = false; uint32_t scan_blocks = 0, flag_cells = 0; } build;
	bool async_build = true;           // builds into tables that already exist skip the host round trip (fr_set_async_build)
	int last_passes = 0;               // passes of the pending render (repeated after a rebuild)
	bool march_timed = false;
	bool zero_counters_in_depth = false;   // this render call: k_depth_clear zeroes the march counters (one memset less)
	// per-stage CUDA events (fr_get_timings).  The lanes of a sequence switch them off: every record is one more
	// command through PCIe per stage, and launch / completion traffic is what the bulk copies of a host -> host
	// sequence slow down (tools/e2e_probe2.py: background copies alone take the device-resident sequence from 0.29
	// to 0.42-0.55 ms per frame)
	bool stage_timing = true;
	int build_timed = 0;               // 1 / 2: grid (and upload) events of a frame build wait to be read

	fr_settings settings{};
	bool have_settings = false;
	fr_camera camera{};
	bool have_camera = false;
	bool have_depth = false;
	int part_rank = 0, part_world = 1, part_tw = 64, part_th = 64;
	int region[4] = { 0, 0, 0, 0 };    // x0, y0, x1, y1 of fr_set_region_partition; x1 == 0: none
	int count_mode = FR_COUNT_CENTRE_BOX;   // reading of find_neighbors_box (fr_set_count_mode)

	std::vector<Frame> frames;
	std::vector<int> pending_frames;   // frames built since the last host wait (resolve_frame at the next one)
	const float* lane_d_xyz = nullptr; size_t lane_n = 0; bool lane_h2d = false;      // a sequence lane's frame between its steps

	// images
	float* d_depth = nullptr;
	float4* d_pos = nullptr;           // where the march writes positions / normals: the context's own images or imported
	float4* d_nrm = nullptr;           // Vulkan memory (fr_import_vk_images_fd)
	float4* d_pos_own = nullptr;
	float4* d_nrm_own = nullptr;
	uchar4* d_rgba = nullptr;
	uchar4* d_rgba_target = nullptr;   // where the colour pass writes (internal or external)
	void* peer_rgba = nullptr;         // another process's colour image opened through CUDA IPC (tile-parallel)
	// scratch for frame builds
	float* d_xyz = nullptr;           size_t cap_xyz = 0;       // staged input particles (n*3)
	unsigned char* h_stage = nullptr; size_t cap_stage = 0;     // pinned: a particle file as read from disk (fm_bgeo.cu)
	uint32_t* d_raw = nullptr;        size_t cap_raw = 0;       // the file's big-endian point block on the device
	uchar4* d_bmp = nullptr;          size_t cap_bmp = 0;       // recording: B G R A rows bottom-up (fm_record.cu)
	uint8_t* h_bmp = nullptr;         size_t cap_bmp_host = 0;  // pinned: header + pixels of one .bmp
	uint32_t* d_keys = nullptr;       size_t cap_keys = 0;
	float4* d_sort_tmp = nullptr;     size_t cap_sort_tmp = 0;     // r = h_ext search: counting sort output before the in-cell ordering
	uint32_t* d_tmp_idx = nullptr;    size_t cap_tmp_idx = 0;      // counting sort output (particle indices) before the in-cell ordering
	uint32_t* d_scan_tmp = nullptr;   size_t cap_scan_tmp = 0;
	float* d_aabb_partial = nullptr;  size_t cap_aabb_partial = 0;  // per-block extrema of k_aabb_params (+ its block ticket)
	float* d_smooth = nullptr;        size_t cap_smooth = 0;       // fr_smooth_depth: smoothed depth, screen normals, kernel weights
	uint32_t* d_tile_bound = nullptr; size_t cap_tile_bound = 0;   // depth pre-pass: per-tile upper bounds
	float* d_splat = nullptr;         size_t cap_splat = 0;        // depth pre-pass: per-particle splat parameters
	uint32_t* d_survivors = nullptr;  size_t cap_survivors = 0;    // depth pre-pass: [0] count, [4..] particle indices
	bool depth_refine_bounds = true;
	uint32_t* d_tiles = nullptr;      size_t cap_tiles = 0;        // march: [0] count, [1] cursor, [2..] covered 8x4 tiles
	int march_ctas_per_sm = 0, march_long_ctas_per_sm = 0, march_coop_ctas_per_sm = 0, march_ctas_per_sm_aniso = 0;
	float4* d_rayq = nullptr;         size_t cap_rayq = 0;         // march: ray queues between the phases
	DeviceCounters* d_counters = nullptr;
	DeviceCounters* h_counters = nullptr;

	fr_timings timings{};
	uint64_t kernel_launches = 0;      // kernels of this library launched since fr_create
	// external interop
	cudaExternalMemory_t ext_mem = nullptr, ext_mem_pos = nullptr, ext_mem_nrm = nullptr;
	cudaExternalSemaphore_t ext_wait = nullptr, ext_signal = nullptr;
};

void set_error(const std::string& msg);
int stream_sync(Context* c);          // waits for everything on c->stream (see Context::blocking_sync)
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define FM_TIME(ctx, event, stream)                                              \
	do { if ((ctx)->stage_timing) FM_CUDA(cudaEventRecord((event), (stream))); } while (0)

// Programmatic dependent launch (sm_90+): the kernels of a frame are queued with the programmatic-stream-serialization
// attribute and begin with fm::pdl_enter() (fm_common.cuh) -- `griddepcontrol.launch_dependents` lets the next kernel's
// CTAs be scheduled as soon as every CTA of this one is resident, `griddepcontrol.wait` holds them until the previous
// grid has completed and its writes are visible.  What overlaps is launch latency and CTA scheduling (2-3 us per kernel
// boundary, 15 kernels per frame); the data dependences are unchanged.  FLUIDMARCH_PDL=0 launches them plainly.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid;
	cfg.blockDim = block;
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at;
	cfg.numAttrs = pdl_enabled() ? 1u : 0u;
	return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#define FM_CUDA(expr)                                                            \
	do {                                                                         \
		cudaError_t _e = (expr);                                                 \
		if (_e != cudaSuccess) return fm::cuda_fail(_e, #expr, __FILE__, __LINE__); \
	} while (0)

template <typename T>
int ensure_capacity(T** ptr, size_t* cap, size_t need)
{
	if (need <= *cap && *ptr) return FR_OK;
	if (*ptr) { cudaFree(*ptr); *ptr = nullptr; *cap = 0; }
	size_t const want = need + need / 4 + 64;     // head-room: a spreading fluid grows its tables frame by frame
	FM_CUDA(cudaMalloc((void**)ptr, want * sizeof(T)));
	*cap = want;
	return FR_OK;
}

// fm_grid.cu
int build_frame(Context* ctx, Frame* f, const float* d_xyz, size_t n, float h, float h_ext_mult, bool allow_async = true);
// first half: everything that may need the host (allocations; without usable tables: bounds + one host wait); second
// half: launches only, so a sequence lane can capture it into a CUDA graph
int build_frame_begin(Context* ctx, Frame* f, const float* d_xyz, size_t n, float h, float h_ext_mult, bool allow_async = true);
int build_frame_finish(Context* ctx);
bool build_will_not_wait(const Context* ctx, const Frame* f, bool allow_async);      // the decision build_frame_begin takes
// the host copy of the frame's grid parameters (waits for the build if it is still queued); FR_RETRIED: the tables
// were too small, the frame has been rebuilt -- work queued behind the first build ran on an unusable frame
constexpr int FR_RETRIED = 1;
int resolve_early(Context* ctx, Frame* f);
int resolve_frame(Context* ctx, Frame* f, bool synced = false);     // synced: the caller has just drained the stream
int build_frame_ext(Context* ctx, Frame* f);      // no-op when already built
FrameView make_view(const Frame& f);              // needs resolve_frame
void free_frame_small(Frame& f);

// fm_depth.cu
// st / raw_xyz: the context's stream and nullptr (particles from the frame's sorted array), or the side stream and the
// packed float3 array the frame is being built from (same image: it is a minimum over all fragments)
int launch_depth_prepass(Context* ctx, const Frame& f, cudaStream_t st, const float* raw_xyz);
// fm_march.cu
int launch_march(Context* ctx, const Frame& f, bool do_march, bool do_shade);
// fm_aniso.cu (the one translation unit compiled with -fmad=false)
struct MarchLaunch;
int launch_march_kernels_aniso(Context* ctx, const MarchLaunch& ml);
int march_occupancy_aniso(int* blocks_per_sm);
int query_aniso(Context* ctx, const Frame& f, const fr_settings& s, const float* points_host, size_t m, float* density,
				float* grad, float* g9);
// fm_context.cu: the two halves of a sequence lane's frame
int lane_frame_begin(fr_context* ctx, const fr_seq_job& job, const char* bgeo_path, bool* stage_a_is_launches_only);
int lane_frame_upload(fr_context* ctx, const fr_seq_job& job);
int lane_frame_stage_a(fr_context* ctx, const fr_seq_job& job);
int lane_frame_resolve(fr_context* ctx, const fr_seq_job& job);
int lane_frame_enqueue(fr_context* ctx, const fr_seq_job& job);
int lane_frame_wait(fr_context* ctx);
int render_depth(fr_context* ctx, int passes, bool again);
int render_resolve(fr_context* ctx, int passes);
int render_march(fr_context* ctx, int passes);
// fm_bgeo.cu
int stage_bgeo(Context* ctx, const char* path, size_t* n_out);
// fm_query.cu
int query_neighbors(Context* ctx, const Frame& f, const float* points_host, size_t m, uint32_t* counts,
					uint32_t* ids, size_t cap, bool ext);
int query_density(Context* ctx, const Frame& f, const float* points_host, size_t m, float* density, float* grad);
int selftest_division(Context* ctx, uint64_t n, uint64_t seed, uint64_t* mismatches);
int measure_l2_bandwidth(Context* ctx, size_t bytes, uint32_t reps, float* gbs);

}  // namespace fm

// the opaque handle of the C ABI
struct fr_context : public fm::Context {};
