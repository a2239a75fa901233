/* fluidmarch.h -- C ABI of the B200-native SPH iso-surface ray-march path (libfluidmarch.so).
 *
 * Plain C, `int` status returns (0 = FR_OK, negative = error; text via fr_last_error()),
 * plain pointers and sizes, no C++ or torch types.  One context per GPU; a context is used from
 * one host thread at a time; all its work is ordered on one CUDA stream.  There is NO CPU
 * fallback: every entry point that computes needs an sm_100 device.
 *
 * The reference (Fruup/bachelor-thesis, paths relative to its root) has no FFI layer; its seam
 * for this path is the C++ class RayMarcher used by AdvancedRenderer::Render.  Each entry point
 * below cites the reference interface it replaces.  A header-compatible `class RayMarcher` over
 * this ABI lives in bachelor-thesis_b200/host/RayMarcher.h; INTEGRATION.md shows the binding.
 */
#ifndef FLUIDMARCH_H
#define FLUIDMARCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FR_ABI_VERSION 4

enum
{
	FR_OK = 0,
	FR_ERR_INVALID = -1,     /* bad argument */
	FR_ERR_CUDA = -2,        /* CUDA runtime error (fr_last_error() has the text) */
	FR_ERR_NO_DEVICE = -3,   /* no usable sm_100 device: the library never falls back to the CPU */
	FR_ERR_STATE = -4,       /* call order violated (e.g. render before upload/camera) */
	FR_ERR_UNSUPPORTED = -5  /* reserved: a feature this build does not have */
};

/* passes of fr_render_async, in path order */
enum
{
	FR_PASS_DEPTH = 1,   /* depth pre-pass: replaces CollectRenderData + DepthRenderPass + depth.vert/frag
	                        (src/app/AdvancedRenderer/AdvancedRenderer.cpp:447-485, DepthRenderPass.cpp:45-87) */
	FR_PASS_MARCH = 2,   /* RayMarcher::PerPixel_Isotropic (src/app/AdvancedRenderer/RayMarcher.cpp:256-344) or, with
	                        fr_settings.enable_anisotropy, PerPixel_Anisotropic (:346-423) */
	FR_PASS_SHADE = 4,   /* CompositionRenderPass + composition.frag:70-122 */
	FR_PASS_ALL = 7
};

typedef struct fr_context fr_context;

/* POD mirror of VisualizationSettings (src/app/AdvancedRenderer/RayMarcher.h:12-26);
 * defaults: src/app/AdvancedRenderer/AdvancedRenderer.cpp:18-28 */
typedef struct fr_settings
{
	int32_t frame;               /* Frame: index of the uploaded frame to render */
	int32_t max_steps;           /* MaxSteps   = 128 */
	float step_size;             /* StepSize   = 0.009 */
	float iso_density;           /* IsoDensity = 1.0 */
	int32_t enable_anisotropy;   /* EnableAnisotropy: 0 = PerPixel_Isotropic, 1 = PerPixel_Anisotropic (reference default) */
	float k_n, k_r, k_s;         /* WPCA eigenvalue clamps 0.5, 2, 2000 (RayMarcher.cpp:227-238) */
	int32_t n_eps;               /* 1 */
	/* additions (0 = reference behaviour) */
	int32_t bisection_steps;     /* >0: refine the hit between the last two samples (north_star item 3);
	                                0 = parity mode, hit = first sample with density >= iso */
	int32_t skip_last_pixel;     /* 1: leave pixel W*H-1 untouched like the reference ThreadPool
	                                (src/app/ThreadPool.cpp:50) */
	int32_t fast_normals;        /* 1: the density-gradient sum of the normal (RayMarcher.cpp:333-338) is evaluated as
	                                sum c_i * r_i with c_i = -+6 sig poly(q) / (|r|^2 h), FMA + approximate reciprocal,
	                                instead of the reference's per-component IEEE divisions (Kernel.cpp:43-51).
	                                Hit mask and positions are unaffected (bit-exact); normals agree to ~1e-6.
	                                0 = normals bit-identical to the reference */
} fr_settings;

/* camera inputs the path reads: CameraController3D::{Position,System} and
 * Camera3D::{View,Projection,InvProjectionView} (src/engine/camera/Camera3D.h:30-32,
 * CameraController3D.h:22-27).  Matrices are column-major, glm memory order m[col*4+row]. */
typedef struct fr_camera
{
	float view[16];
	float projection[16];
	float inv_projection_view[16];
	float position[3];
	float direction[3];          /* CameraController3D::System[2] (CompositionRenderPass.cpp:319) */
} fr_camera;

/* Frame geometry: Frame::{m_Min,m_Max}, DensityGrid::{m_Width,m_Height,m_Depth}
 * (src/app/Dataset.h:27-35,60-62) */
typedef struct fr_frame_info
{
	uint64_t num_particles;
	float h;                     /* Dataset::ParticleRadius (SPH support radius) */
	float min[3], max[3];        /* particle AABB padded by h (Dataset.cpp:78-92) */
	int32_t grid_dims[3];        /* density grid W, H, D (Dataset.cpp:102-104) */
	uint64_t occupied_cells;     /* nodes with Flag set (Dataset.cpp:136-164) */
	int32_t search_min[3];       /* neighbour-search cell range (cells of size h, world origin) */
	int32_t search_dims[3];
} fr_frame_info;

typedef struct fr_counters
{
	uint64_t pixels;             /* W*H */
	uint64_t covered_rays;       /* depth != 1 */
	uint64_t hit_rays;
	uint64_t ray_steps;          /* density evaluations executed */
	uint64_t skip_iterations;    /* empty-cell jumps */
	uint64_t candidates;         /* particles examined (27-cell candidates) */
	uint64_t neighbours;         /* of those, d^2 < h^2 */
	uint64_t early_exits;        /* rays stopped after leaving the grid for good (cannot hit) */
	uint64_t neighbour_overflow; /* samples with more than 8192 neighbours (RayMarcher.cpp:14); must be 0 */
	uint64_t kernel_launches;    /* kernels this context has launched since fr_create (cumulative) */
	uint64_t first_candidates;   /* share of `candidates` examined by k_march_first (first sample of every ray) */
	uint64_t queued_rays;        /* rays that needed more than the first sample (handled by k_march_long) */
	uint64_t first_examined;     /* candidates k_march_first ran the distance test on: its staged walk skips the cells of
	                                the 27 that the search sphere cannot reach (first_candidates counts all 27) */
	uint64_t first_fallbacks;    /* first samples whose tile did not fit the shared-memory stage (walked from global memory) */
} fr_counters;

/* device time of the last call of each stage, milliseconds (CUDA events on the context stream) */
typedef struct fr_timings
{
	float upload_ms;             /* host -> device copy of the particle array */
	float grid_ms;               /* AABB + keys + histogram + scan + scatter + in-cell order + occupancy */
	float depth_ms;
	float march_ms;              /* classify + march + normals + shading (sum of the three below) */
	float download_ms;
	float classify_ms;           /* k_classify: background pixels + tile work list */
	float march_first_ms;        /* k_march_first: first sample of every covered ray (density + gradient), shading */
	float march_long_ms;         /* k_march_long: the rays that need more samples */
} fr_timings;

/* ---- lifetime ------------------------------------------------------------------------------- */
int fr_abi_version(void);
const char* fr_last_error(void);
/* RayMarcher::RayMarcher + the W/H read from Vulkan.SwapchainExtent in Prepare (RayMarcher.cpp:64-69,86-89) */
int fr_create(int device, int width, int height, fr_context** out);
int fr_resize(fr_context* ctx, int width, int height);
/* RayMarcher::Exit (RayMarcher.cpp:71-74) */
void fr_destroy(fr_context* ctx);

/* pinned host memory for uploads/downloads that should overlap with rendering */
int fr_host_alloc(size_t bytes, void** out);
void fr_host_free(void* p);

/* ---- frames: Dataset::ReadFile -> Frames.emplace_back -> Frame::Frame (Dataset.cpp:9-24,292-306) - */
/* xyz: n packed float3 (glm::vec3 AoS as partio delivers them).  Builds, on the device, the
 * neighbour search (Frame::BuildSearch), the AABB (ComputeAABB) and the occupancy grid
 * (BuildDensityGrid).  h = particleRadius, h_ext_mult = particleRadiusMultiplier (assets/config.yml:19-20).
 * Returns once xyz_host has been consumed; the build stays queued on the context's stream.
 * The first build of a frame slot waits once for the device (the table sizes depend on the particle bounds) and
 * reports an unusable frame at once.  Later builds into the same slot reuse its tables and do not wait: the build
 * kernels read the grid parameters from device memory, the first of them also stores them into mapped host memory, and
 * the next fr_render_async of the frame picks them up (its depth pre-pass is queued first, so the GPU never idles) and
 * reports non-finite coordinates (FR_ERR_INVALID) or, if the tables are too small for the new bounds, rebuilds the
 * frame transparently.  (A collapsed simulation -- thousands to millions of particles in one h-cell -- builds too: such a
 * cell is ordered by a sort instead of the quadratic ranking.) */
int fr_upload_frame(fr_context* ctx, int frame, const float* xyz_host, size_t n, float h, float h_ext_mult);
/* 0: every frame build waits for the device once and reports its errors immediately (round-1 behaviour); default 1 */
int fr_set_async_build(fr_context* ctx, int on);
/* 1 (default): CUDA events between the stages of a frame feed fr_get_timings.  0: no events -- the kernels of a frame
 * then form one chain of programmatic dependent launches (each kernel's CTAs are scheduled while the previous kernel
 * drains), and the depth pre-pass of a render queued behind a frame build runs on a second stream beside the build's
 * kernels (from the same particle array: see fr_build_frame_device for how long that array must stay unchanged), the
 * outputs of the uncovered pixels are written on that stream while the long rays march -- the
 * lowest latency for a host that marches one frame at a time as AdvancedRenderer.cpp:257-298 does; fr_get_timings
 * reports zeros.  Sequence lanes always run without the events.  Environment: FLUIDMARCH_OVERLAP=0 keeps everything on
 * the one stream (FLUIDMARCH_BG=0: only the uncovered pixels), FLUIDMARCH_PDL=0 switches the programmatic launches off. */
int fr_set_stage_timing(fr_context* ctx, int on);
/* same, particles already resident in device memory (n packed float3); the build is left on the context's stream:
 * xyz_device must stay valid and unchanged until the host next waits for the context (fr_wait, fr_download, ...) */
int fr_build_frame_device(fr_context* ctx, int frame, const float* xyz_device, size_t n, float h, float h_ext_mult);
/* OctreeNode::NumParticles = |m_Search.find_neighbors_box(cell centre)| (Dataset.cpp:117-131).  find_neighbors_box exists
 * only in the reference's un-vendored CompactNSearch fork, so its reading is a switch; applies to frames built afterwards:
 *   FR_COUNT_CENTRE_BOX (default): particles p with centre - r/2 <= p < centre + r/2 on every axis (FP32), centre = the
 *                                  query point m_Min + (vec3(x,y,z) + 0.5) * cellWidth -- what the reference's own
 *                                  translation units compute under oracle/_ref (oracle/ref/shim/CompactNSearch.h)
 *   FR_COUNT_CELL_EXACT:           particles p with QueryDensityGrid(p) == the node
 * The two differ only for particles within an ulp of a cell face (41 of 1M at BASELINE C2). */
enum { FR_COUNT_CELL_EXACT = 0, FR_COUNT_CENTRE_BOX = 1 };
int fr_set_count_mode(fr_context* ctx, int mode);
int fr_get_frame_info(fr_context* ctx, int frame, fr_frame_info* out);
int fr_release_frame(fr_context* ctx, int frame);
/* parity access to the built structures (any pointer may be NULL):
 *   sorted_xyzi    : n * 4 floats, particles in search order, .w = original index (uint32 bits)
 *   cell_start     : search_dims product + 1 uint32 (exclusive prefix of the per-cell counts)
 *   grid_counts    : grid_dims product uint32, OctreeNode::NumParticles
 *   grid_flags     : grid_dims product uint8, OctreeNode::Flag */
int fr_download_frame(fr_context* ctx, int frame, float* sorted_xyzi, uint32_t* cell_start,
					  uint32_t* grid_counts, uint8_t* grid_flags);

/* ---- per-render state: RayMarcher::Prepare (RayMarcher.cpp:76-100) ------------------------------ */
int fr_set_settings(fr_context* ctx, const fr_settings* s);
int fr_set_camera(fr_context* ctx, const fr_camera* cam);
/* depth supplied by the caller, as Prepare's `float* depth` (W*H floats, row 0 = top, 1.0 = empty);
 * use it together with passes that omit FR_PASS_DEPTH */
int fr_set_depth(fr_context* ctx, const float* depth_host);
/* multi-GPU tile-parallel: this context renders only screen tiles t with t % world == rank
 * (tiles of tile_w x tile_h pixels, row-major tile index); world = 1 renders everything */
int fr_set_tile_partition(fr_context* ctx, int rank, int world, int tile_w, int tile_h);

/* multi-GPU region-parallel: this context renders the pixel rectangle [x0, x1) x [y0, y1) only (bounds multiples of 64
 * or the image edge; all zero = off) AND builds the frames uploaded afterwards from the particles that can influence those
 * pixels under the camera set at build time (fr_set_camera first; a camera change needs a new upload): the grid geometry
 * is the whole frame's, so the pixels are bit-identical to an unpartitioned render, while build, pre-pass and march
 * all shrink with the region.  Images are valid inside the rectangle only; fr_query_* see the kept particles only. */
int fr_set_region_partition(fr_context* ctx, int x0, int y0, int x1, int y1);

/* ---- render: RayMarcher::Start / IsDone (RayMarcher.cpp:102-112, RayMarcher.h:44) ---------------- */
int fr_render_async(fr_context* ctx, int passes);
int fr_is_done(fr_context* ctx);      /* 1 = done, 0 = still running, <0 = error */
int fr_wait(fr_context* ctx);

/* ---- results ---------------------------------------------------------------------------------- */
/* host copies (any pointer may be NULL): depth W*H floats; positions/normals W*H*4 floats
 * (glm::vec4 images of Prepare); rgba W*H*4 bytes, sRGB-encoded R,G,B + linear A.  Waits for the render. */
int fr_download(fr_context* ctx, float* depth, float* positions, float* normals, uint8_t* rgba);
/* device pointers of the same images (valid until fr_resize/fr_destroy), for interop / NCCL */
int fr_device_images(fr_context* ctx, void** depth, void** positions, void** normals, void** rgba);
/* render the colour image into caller-owned device memory instead (e.g. a torch tensor used as
 * an NCCL buffer, or imported Vulkan memory); NULL restores the internal image */
int fr_set_color_target(fr_context* ctx, void* rgba_device);
/* tile-parallel without a gather: the presenting process exports its internal colour image, every other process
 * (one per GPU of the box) opens it as its colour target and renders its tiles (fr_set_tile_partition) straight into
 * the presenter's memory over NVLink / NVSwitch peer access.  The presenter reads the image once all ranks have
 * finished (fr_wait on each + a barrier of the host's choice).  The reference has no counterpart (single GPU). */
typedef struct fr_ipc_handle { unsigned char bytes[64]; } fr_ipc_handle;
int fr_ipc_export_color(fr_context* ctx, fr_ipc_handle* out);
int fr_ipc_open_color_target(fr_context* ctx, const fr_ipc_handle* handle);
int fr_ipc_close_color_target(fr_context* ctx);      /* back to the internal image */
/* generic form of the above for frame rings and flags: plain device allocations (cudaMalloc, hence exportable), their
 * 64-byte handles for the other processes of the box, and the mapping of such a handle into this process */
int fr_device_alloc(int device, size_t bytes, void** out);
int fr_device_free(int device, void* p);
int fr_ipc_export_buffer(int device, void* device_ptr, fr_ipc_handle* out);
int fr_ipc_open_buffer(int device, const fr_ipc_handle* handle, void** out);
int fr_ipc_close_buffer(int device, void* p);
int fr_get_counters(fr_context* ctx, fr_counters* out);
int fr_get_timings(fr_context* ctx, fr_timings* out);
/* the CUDA stream all work of this context is ordered on (a cudaStream_t) */
int fr_get_stream(fr_context* ctx, void** stream);

/* ---- point queries (parity of Dataset::GetNeighbors, Dataset.cpp:272-280) ------------------------- */
/* for each of m query points (packed float3, host): counts[i] = |{j : |x_j - p_i|^2 < h^2}| and, if
 * ids != NULL, up to cap original particle indices at ids[i*cap ...] in the reference's result order */
int fr_query_neighbors(fr_context* ctx, int frame, const float* points_host, size_t m,
					   uint32_t* counts, uint32_t* ids, size_t cap);
/* density = sum_j W(x_j - p_i) (RayMarcher.cpp:322-325) and, if grad != NULL, the un-normalised
 * sum_j gradW(x_j - p_i) (RayMarcher.cpp:333-336), m*3 floats; over the first MAX_NEIGHBORS = 8192 neighbours in the
 * reference's list order, as the march (RayMarcher.cpp:312) */
int fr_query_density(fr_context* ctx, int frame, const float* points_host, size_t m, float* density, float* grad);

/* ---- anisotropic path probes (parity of Dataset::GetNeighborsExt, RayMarcher::WPCA, AnisotropicKernel) ------ */
/* like fr_query_neighbors for the r = h_ext search (Dataset.cpp:282-290): |x_j - p_i|^2 < h_ext^2 */
int fr_query_neighbors_ext(fr_context* ctx, int frame, const float* points_host, size_t m,
						   uint32_t* counts, uint32_t* ids, size_t cap);
/* per point: G = WPCA(p, GetNeighborsExt(p)) with the k_n/k_r/k_s/n_eps of fr_set_settings (g9: m*9 floats, glm::mat3
 * column-major, RayMarcher.cpp:114-254), density = sum W(G, det G, r) over the h-subset (:402-403) and the
 * un-normalised sum of gradW (:411-414, m*3 floats).  grad and g9 may be NULL */
int fr_query_anisotropic(fr_context* ctx, int frame, const float* points_host, size_t m, float* density, float* grad, float* g9);
/* the r = h_ext search structure (Frame::m_SearchExt / m_ParticlesExt, Dataset.cpp:65-75), layout as fr_download_frame;
 * any pointer may be NULL.  Built on first use. */
int fr_download_frame_ext(fr_context* ctx, int frame, float* sorted_xyzi, uint32_t* cell_start, int32_t search_min[3],
						  int32_t search_dims[3]);

/* device self-test of the shortcuts that claim bit-identity with IEEE arithmetic: the shared-reciprocal quotients of
 * gradW (Kernel.cpp:43) on n pseudo-random operand sets, and the unguarded square root / reciprocal sequences of the
 * spline on EVERY float in [2^-100, 2^100]; *mismatches must come back 0 */
int fr_selftest_division(fr_context* ctx, uint64_t n, uint64_t seed, uint64_t* mismatches);
/* roofline denominator the driver does not measure: read-only streaming over `bytes` of device memory that stay
 * resident in the L2 (pick 16-64 MB), `reps` passes, GB/s served by the L2 to the SMs */
int fr_measure_l2_bandwidth(fr_context* ctx, size_t bytes, uint32_t reps, float* gbs);

/* ---- CUDA-Vulkan hand-off: replaces BilateralBuffer::CopyToGPU/CopyFromGPU
 *      (src/app/AdvancedRenderer/BilateralBuffer.cpp:78-134) ------------------------------------------ */
/* imports VkDeviceMemory exported with VK_KHR_external_memory_fd (opaque fd, linear W*H*4-byte
 * image or buffer at `offset`) and makes it the colour target */
int fr_import_vk_memory_fd(fr_context* ctx, int fd, size_t allocation_bytes, size_t offset);
/* the same for the two RGBA32F images the reference's composition pass samples (AdvancedRenderer.cpp:303-304,
 * PositionsBuffer / NormalsBuffer): two exported linear W*H*16-byte buffers become the march's position and normal
 * targets, so that BilateralBuffer::CopyToGPU is a device-local buffer -> image copy instead of a PCIe upload */
int fr_import_vk_images_fd(fr_context* ctx, int positions_fd, int normals_fd, size_t allocation_bytes);
/* VK_KHR_external_semaphore_fd binary semaphores: the render waits on `wait_fd` (image available)
 * before writing and signals `signal_fd` when the image is complete; -1 = none */
int fr_import_vk_semaphores_fd(fr_context* ctx, int wait_fd, int signal_fd);

/* ---- particle files: Dataset::Dataset's file loop, Partio::read and Dataset::ReadFile (src/app/Dataset.cpp:169-227,
 *      292-306; classic .bgeo version 5, vendor/partio/src/io/BGEO.cpp:200-290, optionally gzip'd) ----------------------- */
typedef struct fr_bgeo_info
{
	uint64_t num_particles;
	uint32_t record_words;       /* 32-bit words per point record: x y z w + the file's other point attributes */
	int32_t compressed;          /* 1: the file is a gzip member */
	uint64_t data_offset;        /* byte offset of the point block in the (uncompressed) file */
	uint64_t file_bytes;         /* size on disk */
} fr_bgeo_info;
int fr_bgeo_probe(const char* path, fr_bgeo_info* out);
/* positions ("position" attribute, what Dataset::ReadFile keeps) as packed float3 into host memory; xyz may be NULL
 * to query *n only */
int fr_bgeo_read(const char* path, float* xyz, uint64_t capacity, uint64_t* n);
/* writes a positions-only file, byte for byte what partio's writeBGEO emits for such a particle set */
int fr_bgeo_write(const char* path, const float* xyz, uint64_t n, int compressed);
/* number of files <prefix>1<suffix>, <prefix>2<suffix>, ... that exist, at most count (count < 0: no limit) */
int fr_dataset_count(const char* prefix, const char* suffix, int count);
/* fr_upload_frame straight from a file: the point block goes to the GPU as it lies in the file, the big-endian
 * swap and the record -> packed xyz gather are a kernel in front of the frame build */
int fr_upload_frame_bgeo(fr_context* ctx, int frame, const char* path, float h, float h_ext_mult);

/* ---- recording: Renderer::_Screenshot (src/engine/renderer/Renderer.cpp:326-415: swapchain -> linear image -> host
 *      R/B swizzle -> stbi_write_bmp(name, W, H, 4, data)), once per frame while g_Recording (AdvancedRenderer.cpp:283-297)
 * The colour image of the last render as that .bmp, byte for byte what stb_image_write emits for the same pixels
 * (file header + BITMAPV4HEADER, 32 bpp BI_BITFIELDS, rows bottom-up, B G R A); swizzle + flip run on the device. */
int fr_encode_bmp(fr_context* ctx, uint8_t* out, size_t capacity, size_t* bytes);   /* out may be NULL: *bytes only */
int fr_write_bmp(fr_context* ctx, const char* path);

/* ---- screen-space smoothing: GaussRenderPass (src/app/AdvancedRenderer/GaussRenderPass.cpp:15-66, assets/shaders/advanced/
 *      gauss.frag:28-47; runs every UI frame in the reference) and the Sobel normal of the unprojected smoothed depth
 *      (composition.frag:50-57,87-104, fullscreen.vert:30-42; `#if 0` in the reference) -----------------------------------
 * ComputeGaussKernel: (N+1)^2 weights, out[j + i*(N+1)], 0 <= N <= 31 */
int fr_gauss_kernel(int gauss_n, float* out);
/* the context's depth image (last FR_PASS_DEPTH or fr_set_depth) blurred with that kernel, Spread = 1, clamp-to-edge:
 * W*H floats to smoothed_host; and, if screen_normals_host != NULL, normalize(cross(dx, dy)) of the Sobel differences of
 * smoothedPosition() over the 8 neighbours, W*H*4 floats (w = 1); inv_projection = Camera3D::InvProjection (glm order) */
int fr_smooth_depth(fr_context* ctx, int gauss_n, const float* inv_projection, float* smoothed_host, float* screen_normals_host);

/* ---- frame sequences: the autoplay loop of AdvancedRenderer::Render (src/app/AdvancedRenderer/AdvancedRenderer.cpp:
 *      275-298: wait for the march, then Frame++) with `lanes` frames in flight on one GPU -------------------------------
 * One fr_context (own stream, images, scratch) and one host worker thread per lane; frame k is rendered on lane
 * k % lanes by the same code path as fr_upload_frame -> fr_render_async -> fr_download, so results are bit-identical;
 * kernels and PCIe copies of different frames overlap.  Calls on one fr_sequence come from one host thread. */
typedef struct fr_sequence fr_sequence;

typedef struct fr_seq_job
{
	const float* xyz;            /* n packed float3: Frame::m_Particles as partio delivers them (Dataset.cpp:292-303) */
	uint64_t n;
	float h, h_ext_mult;         /* particleRadius, particleRadiusMultiplier */
	int32_t xyz_on_device;       /* 0: host memory (pinned for overlap), 1: device memory */
	int32_t passes;              /* FR_PASS_* mask, 0 = FR_PASS_ALL */
	/* host outputs (pinned for overlap), any may be NULL: as fr_download */
	float* depth;
	float* positions;
	float* normals;
	uint8_t* rgba;
	const char* bgeo_path;       /* non-NULL: the frame is this particle file (fr_upload_frame_bgeo on the lane's worker:
	                                files of different lanes are read and decoded in parallel); xyz / n are ignored */
	const char* bmp_path;        /* non-NULL: the finished frame is also written there as fr_write_bmp does (recording) */
	/* frame-parallel presentation (multi-GPU): the finished colour image of THIS frame goes straight into
	 * device memory of the caller's choice -- typically a slot of a frame ring on the presenting GPU, opened with
	 * fr_ipc_open_buffer: the shading epilogue's stores travel over NVLink, no gather, no host copy -- and, behind it
	 * on the lane's stream, done_value is stored to *done_flag_device (same kind of memory; NULL = no flag) */
	void* rgba_device;
	uint32_t* done_flag_device;
	uint32_t done_value;
} fr_seq_job;

int fr_seq_create(int device, int width, int height, int lanes, fr_sequence** out);
void fr_seq_destroy(fr_sequence* seq);
int fr_seq_lanes(fr_sequence* seq);
/* how the lane workers wait for the GPU: 0 (default) inside the driver, one busy core per waiting lane -- fastest while
 * (lanes + 1) x processes <= host cores; 1: poll a word in pinned memory and yield the core between polls, for
 * oversubscribed hosts */
int fr_seq_set_yielding(fr_sequence* seq, int on);
/* the context of a lane (counters, timings, frame info of the lane's last frame); do not render on it directly */
int fr_seq_context(fr_sequence* seq, int lane, fr_context** out);
/* applied to every lane (waits for the frames in flight) */
int fr_seq_set_camera(fr_sequence* seq, const fr_camera* cam);
int fr_seq_set_settings(fr_sequence* seq, const fr_settings* s);
/* queues one frame; blocks while the frame's lane is still busy with frame ticket - lanes.  Returns the ticket
 * (0, 1, 2, ...) or a negative error.  The job's buffers must stay valid until the ticket is done */
int64_t fr_seq_submit(fr_sequence* seq, const fr_seq_job* job);
int fr_seq_wait(fr_sequence* seq, int64_t ticket);      /* status of that frame */
int fr_seq_drain(fr_sequence* seq);                     /* waits for everything; first error since the last drain */
/* device time (CUDA events) of everything submitted between begin and end, milliseconds */
int fr_seq_timer_begin(fr_sequence* seq);
int fr_seq_timer_end(fr_sequence* seq, float* ms);

#ifdef __cplusplus
}
#endif
#endif
