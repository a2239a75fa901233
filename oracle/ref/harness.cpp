// C-callable harness around the reference's OWN, UNMODIFIED translation units
// (src/app/AdvancedRenderer/RayMarcher.cpp, src/app/{Dataset,Kernel,ThreadPool}.cpp,
//  src/engine/camera/{Camera3D,CameraController3D}.cpp), compiled where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/libfluidref.so.
//
// TEST INFRASTRUCTURE ONLY: used (a) to pin oracle/fluid_oracle.c, (b) to generate the golden
// fixtures under tests/golden/, (c) as the `cpu_baseline` / `--impl reference` arm of
// bench.py.  Nothing in the product library links or loads this.
//
// The per-pixel functions are private members; the harness reaches them with the usual
// `#define private public` trick (access specifiers do not change the class layout under the
// Itanium ABI) and drives them from its own deterministic parallel-for over ALL W*H pixels,
// because the reference ThreadPool (a) never processes index W*H-1 (ThreadPool.cpp:50) and
// (b) can deadlock in Exit() after a short job.  The reference pool itself can
// still be exercised with ref_march(..., use_ref_pool = 1) for the baseline timing.

#include <engine/hzpch.h>

#define private public
#include "app/AdvancedRenderer/RayMarcher.h"
#undef private

#include "app/Dataset.h"
#include <engine/renderer/Renderer.h>

#include <chrono>
#include <thread>
#include <Partio.h>
// the vendored copy calls MSVC's two-argument-template sprintf_s in its .hdr writer (not on this path)
#define sprintf_s(buf, ...) snprintf(buf, sizeof(buf), __VA_ARGS__)
#define STB_IMAGE_WRITE_IMPLEMENTATION
#define STB_IMAGE_WRITE_STATIC
#include <stb_image_write.h>   // the reference's vendored copy (vendor/stb_image)
#undef sprintf_s

RefExtentStub& RefExtentStub::GetInstance()
{
	static RefExtentStub s;
	return s;
}

// global (non-static) helper defined in the reference's RayMarcher.cpp:51-62
glm::vec3 intersectAABB(glm::vec3 rayOrigin, glm::vec3 rayDir, glm::vec3 boxMin, glm::vec3 boxMax);

namespace
{
struct RefDataset
{
	RefDataset() = default;
	// constructed in place: a Dataset must never be copied or moved once it has frames,
	// the searches hold raw pointers into Frame::m_Particles (Dataset.cpp:204-207)
	RefDataset(const char* prefix, const char* suffix, float h, float mult, int count) :
		ds(prefix, suffix, h, mult, count) {}

	Dataset ds;
	RayMarcher* marcher = nullptr;   // leaked on purpose: ThreadPool::Exit may deadlock
};

double now_s()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// sizeof(ThreadLocals) in RayMarcher.cpp is 2*(4 + 8192*12) = 196616 bytes
constexpr size_t kLocalsBytes = 256 * 1024;
}  // namespace

extern "C" {

int ref_abi_version() { return 4; }

// ---- kernels (Kernel.cpp) ----------------------------------------------------------------
float ref_W(float h, const float* r)
{
	CubicSplineKernel k(h);
	return k.W(glm::vec3(r[0], r[1], r[2]));
}

float ref_W0(float h)
{
	CubicSplineKernel k(h);
	return k.W0();
}

void ref_gradW(float h, const float* r, float* out)
{
	CubicSplineKernel k(h);
	glm::vec3 g = k.gradW(glm::vec3(r[0], r[1], r[2]));
	out[0] = g.x; out[1] = g.y; out[2] = g.z;
}

void ref_intersectAABB(const float* o, const float* d, const float* bmin, const float* bmax, float* out)
{
	glm::vec3 r = intersectAABB(glm::vec3(o[0], o[1], o[2]), glm::vec3(d[0], d[1], d[2]),
								glm::vec3(bmin[0], bmin[1], bmin[2]), glm::vec3(bmax[0], bmax[1], bmax[2]));
	out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

// ---- camera (Camera3D.cpp, CameraController3D.cpp) -----------------------------------------
// All matrices are written column-major (glm memory order), 16 floats each.
void ref_camera(float fov, float aspect, float znear, float zfar, float R, float rotX, float rotY,
				float* view, float* proj, float* inv_proj, float* inv_proj_view,
				float* position, float* system9)
{
	Camera3D cam(fov, aspect, znear, zfar);
	CameraController3D ctl(cam);
	ctl.R = R;
	ctl.RotationX = rotX;
	ctl.RotationY = rotY;
	ctl.ComputeMatrices();
	std::memcpy(view, &cam.View[0][0], 64);
	std::memcpy(proj, &cam.Projection[0][0], 64);
	std::memcpy(inv_proj, &cam.InvProjection[0][0], 64);
	std::memcpy(inv_proj_view, &cam.InvProjectionView[0][0], 64);
	std::memcpy(position, &ctl.Position[0], 12);
	std::memcpy(system9, &ctl.System[0][0], 36);
}

// ---- dataset / frame (Dataset.cpp) ---------------------------------------------------------
void* ref_dataset_create(const float* xyz, size_t n, float h, float mult, int box_mode, double* build_seconds)
{
	CompactNSearch::box_mode() = box_mode;
	auto* r = new RefDataset();
	Dataset& ds = r->ds;
	// same field set-up as Dataset::makeCube (Dataset.cpp:235-240)
	ds.ParticleRadius = h;
	ds.ParticleRadiusExt = mult * h;
	ds.ParticleRadiusInv = 1.0f / h;
	ds.ParticleRadiusExtInv = 1.0f / ds.ParticleRadiusExt;
	ds.m_IsotropicKernel = CubicSplineKernel(ds.ParticleRadius);
	ds.m_AnisotropicKernel = AnisotropicKernel(ds.ParticleRadius);

	std::vector<Particle> particles(n);
	std::memcpy(particles.data(), xyz, n * 12);

	ds.Frames.reserve(1);   // frames must never move (Dataset.cpp:204-207)
	double const t0 = now_s();
	ds.Frames.emplace_back(&ds, particles, ds.ParticleRadius, ds.ParticleRadiusExt);
	double const t1 = now_s();
	if (build_seconds) *build_seconds = t1 - t0;
	ds.MaxParticles = n;
	ds.Loaded = true;
	return r;
}

// Loads a sequence through the reference's own partio path (Dataset.cpp:169-227).
void* ref_dataset_load(const char* prefix, const char* suffix, float h, float mult, int count, int box_mode)
{
	CompactNSearch::box_mode() = box_mode;
	auto* r = new RefDataset(prefix, suffix, h, mult, count);
	return r;
}

void ref_dataset_destroy(void* p)
{
	delete static_cast<RefDataset*>(p);
}

int ref_num_frames(void* p) { return int(static_cast<RefDataset*>(p)->ds.Frames.size()); }

size_t ref_frame_num_particles(void* p, int f) { return static_cast<RefDataset*>(p)->ds.Frames[f].m_Particles.size(); }

void ref_frame_info(void* p, int f, float* mn, float* mx, int* dims)
{
	Frame& fr = static_cast<RefDataset*>(p)->ds.Frames[f];
	std::memcpy(mn, &fr.m_Min[0], 12);
	std::memcpy(mx, &fr.m_Max[0], 12);
	dims[0] = int(fr.m_DensityGrid.m_Width);
	dims[1] = int(fr.m_DensityGrid.m_Height);
	dims[2] = int(fr.m_DensityGrid.m_Depth);
}

// particles after the in-place Morton permutation (Dataset.cpp:58-62)
void ref_frame_particles(void* p, int f, float* out_xyz)
{
	Frame& fr = static_cast<RefDataset*>(p)->ds.Frames[f];
	std::memcpy(out_xyz, fr.m_Particles.data(), fr.m_Particles.size() * 12);
}

void ref_frame_grid(void* p, int f, uint32_t* num_particles, uint8_t* flags)
{
	Frame& fr = static_cast<RefDataset*>(p)->ds.Frames[f];
	auto const& nodes = fr.m_DensityGrid.m_Nodes;
	for (size_t i = 0; i < nodes.size(); i++)
	{
		if (num_particles) num_particles[i] = nodes[i].NumParticles;
		if (flags) flags[i] = nodes[i].Flag ? 1 : 0;
	}
}

// index of the density-grid cell a point falls in, -1 outside (Frame::QueryDensityGrid)
int64_t ref_query_cell(void* p, int f, const float* x)
{
	Frame& fr = static_cast<RefDataset*>(p)->ds.Frames[f];
	OctreeNode* n = fr.QueryDensityGrid(glm::vec3(x[0], x[1], x[2]));
	return n ? int64_t(n - fr.m_DensityGrid.m_Nodes.data()) : -1;
}

// Dataset::GetNeighbors / GetNeighborsExt; returns the count, writes up to cap indices
size_t ref_neighbors(void* p, int f, const float* x, int ext, uint32_t* out, size_t cap)
{
	Dataset& ds = static_cast<RefDataset*>(p)->ds;
	glm::vec3 q(x[0], x[1], x[2]);
	std::vector<uint32_t> nb = ext ? ds.GetNeighborsExt(q, uint32_t(f)) : ds.GetNeighbors(q, uint32_t(f));
	for (size_t i = 0; i < nb.size() && i < cap; i++) out[i] = nb[i];
	return nb.size();
}

// ---- the march (RayMarcher.cpp) --------------------------------------------------------------
// positions/normals: W*H*4 floats, depth: W*H floats, row-major, row 0 = top.
// inv_proj_view: 16 floats column-major; cam_pos: 3 floats.
// threads <= 0 -> hardware_concurrency().  use_ref_pool: 1 = run through the reference's own
// ThreadPool (hardware_concurrency()-1 workers; last pixel not processed), 0 = harness
// parallel-for over all pixels.  Returns elapsed wall seconds of the march.
double ref_march(void* p, int frame, int W, int H,
				 int max_steps, float step_size, float iso, int anisotropic,
				 float k_n, float k_r, float k_s, int N_eps,
				 const float* inv_proj_view, const float* cam_pos,
				 const float* depth, float* positions, float* normals,
				 int threads, int use_ref_pool)
{
	RefDataset* r = static_cast<RefDataset*>(p);
	if (!r->marcher) r->marcher = new RayMarcher();

	Vulkan.SwapchainExtent.width = uint32_t(W);
	Vulkan.SwapchainExtent.height = uint32_t(H);

	VisualizationSettings s{};
	s.Frame = frame;
	s.MaxSteps = max_steps;
	s.StepSize = step_size;
	s.IsoDensity = iso;
	s.EnableAnisotropy = anisotropic != 0;
	s.k_n = k_n; s.k_r = k_r; s.k_s = k_s; s.N_eps = N_eps;

	Camera3D cam(glm::radians(60.0f), float(W) / float(H), 0.1f, 1000.0f);
	CameraController3D ctl(cam);
	std::memcpy(&cam.InvProjectionView[0][0], inv_proj_view, 64);
	std::memcpy(&ctl.Position[0], cam_pos, 12);

	RayMarcher& m = *r->marcher;
	m.Prepare(s, ctl, &r->ds, reinterpret_cast<glm::vec4*>(positions),
			  reinterpret_cast<glm::vec4*>(normals), const_cast<float*>(depth));

	double const t0 = now_s();
	if (use_ref_pool)
	{
		m.Start();
		while (!m.IsDone()) std::this_thread::sleep_for(std::chrono::microseconds(200));
		double const t1 = now_s();
		// IsDone() can be true while workers finish their last pixel (ThreadPool.cpp:33-36)
		std::this_thread::sleep_for(std::chrono::milliseconds(100));
		return t1 - t0;
	}

	uint32_t const total = uint32_t(W) * uint32_t(H);
	int nt = threads > 0 ? threads : int(std::thread::hardware_concurrency());
	if (nt < 1) nt = 1;
	std::atomic<uint32_t> next{ 0 };
	constexpr uint32_t kChunk = 64;
	auto work = [&]() {
		void* locals = std::malloc(kLocalsBytes);
		for (;;)
		{
			uint32_t const b = next.fetch_add(kChunk);
			if (b >= total) break;
			uint32_t const e = std::min(total, b + kChunk);
			for (uint32_t i = b; i < e; i++)
			{
				if (anisotropic) m.PerPixel_Anisotropic(i, locals);
				else m.PerPixel_Isotropic(i, locals);
			}
		}
		std::free(locals);
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < nt; t++) pool.emplace_back(work);
	work();
	for (auto& t : pool) t.join();
	return now_s() - t0;
}

// ---- anisotropic path probes (RayMarcher.cpp:114-254, Kernel.cpp:55-125) ----------------------------
// Eigen::SelfAdjointEigenSolver<Matrix3f>::computeDirect on a symmetric matrix given column-major (9 floats);
// only the lower triangle is read, as in RayMarcher::WPCA (RayMarcher.cpp:179-181).
int ref_eigen3(const float* c9, float* evals3, float* evecs9)
{
	Eigen::Matrix3f C;
	for (int c = 0; c < 3; c++)
		for (int r = 0; r < 3; r++) C(r, c) = c9[3 * c + r];
	Eigen::SelfAdjointEigenSolver<Eigen::Matrix3f> solver;
	solver.computeDirect(C);
	Eigen::Vector3f const S = solver.eigenvalues();
	Eigen::Matrix3f const R = solver.eigenvectors();
	for (int k = 0; k < 3; k++) evals3[k] = S[k];
	for (int c = 0; c < 3; c++)
		for (int r = 0; r < 3; r++) evecs9[3 * c + r] = R(r, c);
	return solver.info() == Eigen::Success ? 0 : 1;
}

// RayMarcher::WPCA itself.  nbr_xyz: N absolute neighbour positions in list order.  g9: glm::mat3 G, column-major.
void ref_wpca(void* p, float k_n, float k_r, float k_s, int N_eps, const float* particle, const float* nbr_xyz,
			  uint32_t N, float* g9)
{
	RefDataset* r = static_cast<RefDataset*>(p);
	if (!r->marcher) r->marcher = new RayMarcher();
	RayMarcher& m = *r->marcher;
	m.m_Dataset = &r->ds;
	m.m_Settings.k_n = k_n; m.m_Settings.k_r = k_r; m.m_Settings.k_s = k_s; m.m_Settings.N_eps = N_eps;
	glm::mat3 G(0.0f);
	m.WPCA(glm::vec3(particle[0], particle[1], particle[2]), reinterpret_cast<const glm::vec3*>(nbr_xyz), N, G, false);
	std::memcpy(g9, &G[0][0], 36);
}

float ref_det3(const float* g9)
{
	glm::mat3 G;
	std::memcpy(&G[0][0], g9, 36);
	return glm::determinant(G);
}

float ref_aniso_W(float h, const float* g9, float detG, const float* r)
{
	AnisotropicKernel k(h);
	glm::mat3 G;
	std::memcpy(&G[0][0], g9, 36);
	return k.W(G, detG, glm::vec3(r[0], r[1], r[2]));
}

void ref_aniso_gradW(float h, const float* g9, float detG, const float* r, float* out)
{
	AnisotropicKernel k(h);
	glm::mat3 G;
	std::memcpy(&G[0][0], g9, 36);
	glm::vec3 g = k.gradW(G, detG, glm::vec3(r[0], r[1], r[2]));
	out[0] = g.x; out[1] = g.y; out[2] = g.z;
}

float ref_cubic_W(float h, const float* r)
{
	CubicKernel k(h);
	return k.W(glm::vec3(r[0], r[1], r[2]));
}

// particles of the r = h_ext search after ITS Morton permutation (Frame::m_ParticlesExt, Dataset.cpp:65-75)
void ref_frame_particles_ext(void* p, int f, float* out_xyz)
{
	Frame& fr = static_cast<RefDataset*>(p)->ds.Frames[f];
	std::memcpy(out_xyz, fr.m_ParticlesExt.data(), fr.m_ParticlesExt.size() * 12);
}

int ref_hardware_threads() { return int(std::thread::hardware_concurrency()); }

// ---- screenshot file: the save step of Renderer::_Screenshot (Renderer.cpp:400-409) after its R/B swizzle, i.e.
// stbi_write_bmp(filename, W, H, 4, rgba) with the reference's vendored stb_image_write
int ref_write_bmp(const char* path, int w, int h, const unsigned char* rgba)
{
	return stbi_write_bmp(path, w, h, 4, rgba);
}

// ---- particle files through the reference's vendored partio (Dataset.cpp:209-220, 292-306) -------------------------
// positions of a file exactly as Dataset::ReadFile takes them (before Frame::BuildSearch permutes them); returns the
// particle count or -1
long ref_partio_read(const char* path, float* out_xyz, size_t cap)
{
	Partio::ParticlesDataMutable* file = Partio::read(path);
	if (!file) return -1;
	Partio::ParticleAttribute attr;
	file->attributeInfo("position", attr);
	long const n = file->numParticles();
	for (long i = 0; i < n && size_t(i) < cap; i++)
		std::memcpy(out_xyz + 3 * i, file->data<float>(attr, int(i)), 12);
	file->release();
	return n;
}

// writes a file with partio's own writer.  extras: bit 0 = "velocity" (vector 3), bit 1 = "id" (int 1),
// bit 2 = "density" (float 1) in front of everything else, bit 3 = an indexed-string attribute "kind"
int ref_partio_write(const char* path, const float* xyz, size_t n, int extras, int compressed)
{
	Partio::ParticlesDataMutable* p = Partio::create();
	Partio::ParticleAttribute dens, pos, vel, id, kind;
	if (extras & 4) dens = p->addAttribute("density", Partio::FLOAT, 1);
	pos = p->addAttribute("position", Partio::VECTOR, 3);
	if (extras & 1) vel = p->addAttribute("velocity", Partio::VECTOR, 3);
	if (extras & 2) id = p->addAttribute("id", Partio::INT, 1);
	if (extras & 8)
	{
		kind = p->addAttribute("kind", Partio::INDEXEDSTR, 1);
		p->registerIndexedStr(kind, "fluid");
		p->registerIndexedStr(kind, "boundary-particle");
	}
	p->addParticles(int(n));
	for (size_t i = 0; i < n; i++)
	{
		std::memcpy(p->dataWrite<float>(pos, int(i)), xyz + 3 * i, 12);
		if (extras & 1) { float* v = p->dataWrite<float>(vel, int(i)); v[0] = float(i); v[1] = -1.5f; v[2] = xyz[3 * i] * 2.0f; }
		if (extras & 2) p->dataWrite<int>(id, int(i))[0] = int(i) * 7 + 1;
		if (extras & 4) p->dataWrite<float>(dens, int(i))[0] = 1000.0f + float(i);
		if (extras & 8) p->dataWrite<int>(kind, int(i))[0] = int(i & 1);
	}
	Partio::write(path, *p, compressed != 0);
	p->release();
	return 0;
}

}  // extern "C"
