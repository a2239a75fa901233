/* fluid_oracle.h -- CPU parity oracle for the SPH iso-surface ray-march hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the checker.
 * The product library (libfluidmarch.so) never links, loads or falls back to it.
 *
 * Plain-C restatement of the reference's algorithm; every function cites the reference
 * file:line it follows (paths relative to /root/reference).  Pinning status: the reference has
 * no tests or golden vectors for this path, so this oracle is pinned against
 * OUTPUTS OF THE REFERENCE ITSELF: oracle/_ref/libfluidref.so is built from the reference's
 * unmodified translation units and tests/test_oracle_vs_ref.py + tests/golden/ check this
 * file against it bit for bit.  One dependency stays unpinned: the reference's neighbour
 * search is an un-vendored fork of CompactNSearch (pinned commit unknown); both this file
 * and the _ref build restate its published algorithm (see oracle/ref/shim/CompactNSearch.h).
 */
#ifndef FLUID_ORACLE_H
#define FLUID_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fo_frame fo_frame;

/* mirror of VisualizationSettings (RayMarcher.h:12-26) without the frame index */
typedef struct fo_settings
{
	int32_t max_steps;      /* MaxSteps, default 128 (AdvancedRenderer.cpp:19) */
	float step_size;        /* StepSize, default 0.009 */
	float iso_density;      /* IsoDensity, default 1.0 */
	int32_t anisotropic;    /* EnableAnisotropy */
	float k_n, k_r, k_s;    /* WPCA eigenvalue clamps */
	int32_t n_eps;
} fo_settings;

typedef struct fo_counters
{
	uint64_t pixels;            /* W*H */
	uint64_t covered_rays;      /* depth != 1 */
	uint64_t hit_rays;
	uint64_t ray_steps;         /* density evaluations */
	uint64_t skip_iterations;   /* empty-cell jumps */
	uint64_t candidates;        /* particles examined by the 27-cell queries */
	uint64_t neighbours;        /* of those, with d^2 < h^2 */
	uint64_t steps_outside_grid; /* ray_steps whose sample lies outside the density grid */
} fo_counters;

int fo_abi_version(void);
void fo_set_threads(int threads);   /* 0 = all online cores (default) */
int fo_get_threads(void);

/* ---- Kernel.cpp -------------------------------------------------------------------- */
float fo_W0(float h);                                       /* Kernel.cpp:8-14, Kernel.h:12 */
float fo_W(float h, const float r[3]);                      /* Kernel.cpp:16-32 */
void fo_gradW(float h, const float r[3], float out[3]);     /* Kernel.cpp:34-52 */
/* RayMarcher.cpp:51-62 */
void fo_intersect_aabb(const float o[3], const float d[3], const float bmin[3], const float bmax[3], float out[3]);
/* deterministic cos(pi/2 * s), s in [0,1]: the depth impostor profile (depth.frag:25) */
float fo_cos_half_pi(float s);

/* ---- Dataset.cpp: Frame ------------------------------------------------------------ */
/* How OctreeNode::NumParticles is counted, i.e. the reading of the fork-only find_neighbors_box
 * (Dataset.cpp:128).  Both count "the particles inside the cell"; they differ only for
 * particles within an ulp of a cell face.
 *   FO_COUNT_CELL_EXACT  a particle counts for the node QueryDensityGrid(particle) returns (the convention
 *                        the CUDA path implements; default)
 *   FO_COUNT_CENTRE_BOX  half-open FP32 box [c - h/2, c + h/2) around the cell centre the reference queries --
 *                        what the stand-in search of the oracle/_ref build does; used to pin this file
 *                        against oracle/_ref bit for bit */
enum { FO_COUNT_CELL_EXACT = 0, FO_COUNT_CENTRE_BOX = 1 };
void fo_set_count_mode(int mode);   /* applies to frames created afterwards */
/* Frame::Frame (Dataset.cpp:9-24): BuildSearch (r = h and r = h_mult*h) + ComputeAABB + BuildDensityGrid */
fo_frame* fo_frame_create(const float* xyz, size_t n, float h, float h_mult);
void fo_frame_destroy(fo_frame* f);
size_t fo_frame_num_particles(const fo_frame* f);
void fo_frame_info(const fo_frame* f, float mn[3], float mx[3], int32_t dims[3]);
void fo_frame_particles(const fo_frame* f, float* xyz);     /* after the Morton permutation */
void fo_frame_grid(const fo_frame* f, uint32_t* counts, uint8_t* flags);
int64_t fo_query_cell(const fo_frame* f, const float p[3]); /* Frame::QueryDensityGrid, -1 outside */
/* Dataset::GetNeighbors / GetNeighborsExt (Dataset.cpp:272-290); returns the count, writes <= cap ids */
size_t fo_neighbors(const fo_frame* f, const float p[3], int ext, uint32_t* out, size_t cap);

/* ---- depth pre-pass (AdvancedRenderer.cpp:447-485, DepthRenderPass.cpp:45-87, depth.vert/frag) */
/* view/proj: 16 floats column-major (glm order).  depth: W*H floats, 1.0 = empty. Returns 0 or <0. */
int fo_depth_prepass(const fo_frame* f, int32_t W, int32_t H, const float view[16], const float proj[16], float* depth);

/* ---- the march (RayMarcher.cpp:256-344) --------------------------------------------- */
/* pos4/nrm4: W*H*4 floats.  band (optional): per pixel min over evaluated samples of
 * |density - iso| (+inf if none).  steps (optional): density evaluations per pixel.
 * first_pixel/num_pixels select a pixel-index range (the reference pool never runs index W*H-1). */
int fo_march(const fo_frame* f, int32_t W, int32_t H, const fo_settings* s,
			 const float inv_proj_view[16], const float cam_pos[3], const float* depth,
			 float* pos4, float* nrm4, float* band, uint32_t* steps, fo_counters* counters,
			 int threads);

/* ---- anisotropic path (RayMarcher.cpp:114-254,346-423; Kernel.cpp:55-125); fluid_oracle_aniso.inc ------ */
/* fo_march runs PerPixel_Anisotropic when fo_settings.anisotropic != 0 */
void fo_frame_particles_ext(const fo_frame* f, float* xyz);     /* m_ParticlesExt after its Morton permutation */
/* Eigen::SelfAdjointEigenSolver<Matrix3f>::computeDirect; c9/evecs9 column-major, lower triangle of c9 read */
int fo_eigen3(const float c9[9], float evals3[3], float evecs9[9]);
/* RayMarcher::WPCA; nbr_xyz = N absolute neighbour positions in list order; g9 = glm::mat3 G, column-major */
void fo_wpca(float h, float h_ext, const fo_settings* s, const float particle[3], const float* nbr_xyz, uint32_t N, float g9[9]);
float fo_det3(const float g9[9]);                               /* glm::determinant(mat3) */
float fo_aniso_W(float h, const float g9[9], float detG, const float r[3]);               /* Kernel.cpp:63-82 */
void fo_aniso_gradW(float h, const float g9[9], float detG, const float r[3], float out[3]); /* Kernel.cpp:84-107 */
float fo_cubic_W(float h, const float r[3]);                    /* CubicKernel::W, Kernel.cpp:117-125 */
/* glibc 2.39 atan2f / sinf / cosf restated op for op (what computeDirect's std::atan2/cos/sin resolve to in the
 * oracle/_ref build); fo_set_trig_libm(1) makes the eigen solver call libm itself instead */
float fo_atan2f(float y, float x);
float fo_sinf(float x);
float fo_cosf(float x);
void fo_set_trig_libm(int on);
/* brute-force comparison of the three restatements with libm: returns the number of mismatching results.
 * which: 0 sinf, 1 cosf over every float in [lo, hi]; 2 atan2f over `count` pseudo-random (y >= 0, x) pairs */
uint64_t fo_trig_selftest(int which, float lo, float hi, uint64_t count, uint32_t seed);

/* ---- shading (composition.frag:37-66,70-122, CompositionRenderPass.cpp:313-321) ------- */
/* color4 (optional): linear RGBA floats.  rgba8 (optional): sRGB-encoded bytes in R,G,B,A order
 * (alpha linear), what a B8G8R8A8_SRGB / R8G8B8A8_SRGB attachment stores. */
void fo_shade(int32_t W, int32_t H, const float* pos4, const float* nrm4,
			  const float inv_proj_view[16], const float cam_pos[3], const float cam_dir[3],
			  float* color4, uint8_t* rgba8);

/* ---- screen-space smoothing (runs in the reference, its consumer in composition.frag is `#if 0`) --------
 * GaussRenderPass.cpp:15-66: kernel[j + i*(N+1)] = e^(-r^2/2) / sum over the (2N+1)^2 taps, r = sqrt(i^2 + j^2).
 * gauss.frag:28-47: SmoothedDepth = sum over i, j in [-N, N] (i outer, j inner) of kernel[|i|*(N+1) + |j|] *
 * Depth(x + i, y + j), clamp-to-edge (BilateralBuffer.cpp:248-250), TexelWidth = Spread / W with Spread = 1: exact texels.
 * composition.frag:50-57,87-104 + fullscreen.vert:30-42: smoothedPosition(uv) = unproject(InvProjection, 2uv - 1,
 * SmoothedDepth(uv)); dx, dy = Sobel over the 8 neighbours (one texel away); screenNormal = normalize(cross(dx, dy)).
 * FP32, one rounding per operation (GLSL may contract a*b+c: unpinned, like the rest of the GLSL restatements). */
void fo_gauss_kernel(int32_t n, float* out);      /* (n+1)^2 floats */
void fo_gauss_depth(int32_t W, int32_t H, const float* depth, int32_t n, float* smoothed);
void fo_sobel_normals(int32_t W, int32_t H, const float* smoothed, const float inv_proj[16], float* nrm4);

#ifdef __cplusplus
}
#endif
#endif
