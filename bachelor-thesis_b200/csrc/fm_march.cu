// fm_march.cu -- the ray-march kernel with fused normals and shading (sm_100a).
//
// Replaces RayMarcher::PerPixel_Isotropic (src/app/AdvancedRenderer/RayMarcher.cpp:256-344) together with
// its callees Frame::QueryDensityGrid (src/app/Dataset.cpp:26-47), Dataset::GetNeighbors (:272-280),
// CubicSplineKernel::W/gradW (src/app/Kernel.cpp:16-52), intersectAABB (RayMarcher.cpp:51-62), and the
// fullscreen composition pass (assets/shaders/advanced/composition.frag:37-66,70-122,
// CompositionRenderPass.cpp:313-321), which only ever reads its own pixel and is therefore fused as the
// epilogue of the same thread.
//
// Mapping: one warp = one 8x4 pixel tile (lanes = rays), 8 warps per CTA = a 32x8 pixel block.  The rays
// of a tile are ~1 cell apart (pixel footprint at the default camera distance ~ h/10), so the 9
// contiguous particle ranges each lane walks are the same addresses across the warp: every LDG.128 of a
// candidate is a single-sector broadcast served by L1/L2 (the sorted particle array, 16 B/particle,
// lives in L2: 16 MB at 1M particles).  No neighbour list is materialised; the d^2 < h^2 test and the
// kernel sum run inline in the reference's accumulation order (see fm_common.cuh: FrameView).
#include "fm_internal.h"

namespace fm
{

namespace
{

constexpr int kMaxNeighbors = 2 * 4096;   // MAX_NEIGHBORS (RayMarcher.cpp:14)

struct MarchParams
{
	int W, H;
	float two_w_inv, two_h_inv;       // m_TwoWidthInv, m_TwoHeightInv (RayMarcher.cpp:88-89)
	float inv_w, inv_h;
	float ipv[16];                    // m_InvProjectionView
	float cam[3];                     // m_CameraPosition
	float dir[3];                     // Uniforms.CameraDirection
	int max_steps;
	float step_size, iso;
	int bisection_steps;
	int skip_last_pixel;
	int early_out;
	int part_rank, part_world, part_tw, part_th, part_tiles_x;
	int do_march, do_shade;
};

struct LaneCounters
{
	uint32_t covered, hits, steps, skips, candidates, neighbours, early_exits, overflow;
};

// density (and optionally the un-normalised gradient sum) at p: Dataset::GetNeighbors + the W / gradW
// loops of RayMarcher.cpp:309-336, fused.  Accumulation order == the reference's neighbour order.
template <bool GRAD>
__device__ __forceinline__ float eval_density(const FrameView& f, f3 p, f3& grad, LaneCounters& lc)
{
	int const kx = search_cell_of(f.search_inv, p.x) - f.kmin.x;
	int const ky = search_cell_of(f.search_inv, p.y) - f.kmin.y;
	int const kz = search_cell_of(f.search_inv, p.z) - f.kmin.z;
	int const z0 = max(kz - 1, 0), z1 = min(kz + 1, f.kdim.z - 1);
	float density = 0.0f;
	f3 g = mk3(0.0f, 0.0f, 0.0f);
	uint32_t nn = 0;
	if (z0 <= z1)
	{
#pragma unroll 1
		for (int dx = -1; dx <= 1; dx++)
		{
			int const x = kx + dx;
			if ((unsigned)x >= (unsigned)f.kdim.x) continue;
#pragma unroll 1
			for (int dy = -1; dy <= 1; dy++)
			{
				int const y = ky + dy;
				if ((unsigned)y >= (unsigned)f.kdim.y) continue;
				uint32_t const base = ((uint32_t)x * (uint32_t)f.kdim.y + (uint32_t)y) * (uint32_t)f.kdim.z;
				uint32_t const b = __ldg(f.cell_start + base + z0);
				uint32_t const e = __ldg(f.cell_start + base + z1 + 1);
				lc.candidates += e - b;
#pragma unroll 4
				for (uint32_t j = b; j < e; j++)
				{
					float4 const q = __ldg(f.sorted + j);
					// CompactNSearch: d = x - xb; l2 = d0*d0 + d1*d1 + d2*d2 (left to right); l2 < r2
					float const d0 = subr(p.x, q.x), d1 = subr(p.y, q.y), d2 = subr(p.z, q.z);
					float const l2 = addr(addr(mulr(d0, d0), mulr(d1, d1)), mulr(d2, d2));
					if (l2 < f.kernel.h_squared)
					{
						if (nn < (uint32_t)kMaxNeighbors)   // list truncation of RayMarcher.cpp:312
						{
							if (GRAD) g = add3(g, spline_gradW_inrange(f.kernel, mk3(-d0, -d1, -d2), l2));
							else density = addr(density, spline_W_inrange(f.kernel, l2));
						}
						nn++;
					}
				}
			}
		}
	}
	lc.neighbours += nn;
	if (nn > (uint32_t)kMaxNeighbors) lc.overflow++;
	grad = g;
	return density;
}

// sampleFloor (composition.frag:37-46)
__device__ __forceinline__ void sample_floor(f3 a, f3 r, float out[4])
{
	float const sN = subr(-1.0f, a.y);                         // FLOOR_HEIGHT - a.y
	f3 const b = add3(a, divs3(scale3(r, sN), r.y));           // a + r * (..) / r.y, left to right
	float const mx = subr(b.x, mulr(2.0f, floorf(divr(b.x, 2.0f))));   // mod(b.x, 2)
	float const mz = subr(b.z, mulr(2.0f, floorf(divr(b.z, 2.0f))));
	float const fx = (1.0f < mx) ? 0.0f : 1.0f;                // step(m, 1)
	float const fy = (1.0f < mz) ? 0.0f : 1.0f;
	float const g = addr(0.25f, divr(addr(fx, fy), 4.0f));
	out[0] = g; out[1] = g; out[2] = g; out[3] = 0.5f;
}

__device__ __forceinline__ uint32_t unorm8(float x)
{
	if (!(x > 0.0f)) return 0u;      // also NaN
	if (x >= 1.0f) return 255u;
	return (uint32_t)(addr(mulr(x, 255.0f), 0.5f));
}

// linear -> sRGB, applied by the B8G8R8A8_SRGB swapchain attachment on write (RendererInit2.cpp:50)
__device__ __forceinline__ float srgb_encode(float c)
{
	if (!(c > 0.0f)) return 0.0f;
	if (c >= 1.0f) return 1.0f;
	return c <= 0.0031308f ? mulr(12.92f, c) : subr(mulr(1.055f, powf(c, 1.0f / 2.4f)), 0.055f);
}

// composition.frag:70-122 for one pixel
__device__ __forceinline__ uchar4 shade_pixel(const MarchParams& mp, int px, int py, float4 P, float4 N)
{
	float const u = divr(addr((float)px, 0.5f), (float)mp.W);   // fullscreen.vert UV at the pixel centre
	float const v = divr(addr((float)py, 0.5f), (float)mp.H);
	// viewRay() (composition.frag:59-66): a far-plane POINT used as a direction
	float wh[4];
	mat4_mul_vec4(mp.ipv, subr(mulr(2.0f, u), 1.0f), subr(mulr(2.0f, v), 1.0f), 1.0f, 1.0f, wh);
	f3 const view_ray = divs3(mk3(wh[0], wh[1], wh[2]), wh[3]);
	f3 const cam = mk3(mp.cam[0], mp.cam[1], mp.cam[2]);
	float color[4];
	if (P.w == 0.0f)
	{
		float fl[4];
		sample_floor(cam, view_ray, fl);
#pragma unroll
		for (int k = 0; k < 4; k++) color[k] = mulr(0.75f, fl[k]);
	}
	else
	{
		f3 const world = mk3(P.x, P.y, P.z);
		f3 const normal = mk3(N.x, N.y, N.z);
		// refract(I, N, eta): k = 1 - eta^2 (1 - dot(N,I)^2); k < 0 ? 0 : eta*I - (eta*dot(N,I) + sqrt(k))*N
		f3 const I = normalize3(view_ray);
		float const eta = 1.333f;
		float const dotNI = dot3(normal, I);
		float const k = subr(1.0f, mulr(mulr(eta, eta), subr(1.0f, mulr(dotNI, dotNI))));
		f3 refracted = mk3(0.0f, 0.0f, 0.0f);
		if (k >= 0.0f) refracted = sub3(scale3(I, eta), scale3(normal, addr(mulr(eta, dotNI), sqrtr(k))));
		float fl[4];
		sample_floor(world, refracted, fl);
		float const fv = -dot3(mk3(mp.dir[0], mp.dir[1], mp.dir[2]), normal);
		float const amb = subr(addr(0.15f, 1.0f), mulr(fv, fv));
		float const diffuse[4] = { 120.0f / 255.0f, 185.0f / 255.0f, 255.0f / 255.0f, 255.0f / 255.0f };
#pragma unroll
		for (int kk = 0; kk < 4; kk++) color[kk] = addr(mulr(fv, fl[kk]), mulr(amb, diffuse[kk]));
	}
	return make_uchar4((unsigned char)unorm8(srgb_encode(color[0])), (unsigned char)unorm8(srgb_encode(color[1])),
					   (unsigned char)unorm8(srgb_encode(color[2])), (unsigned char)unorm8(color[3]));
}

__global__ void __launch_bounds__(256, 3) k_march_shade(FrameView f, MarchParams mp, const float* __restrict__ depth,
													 float4* __restrict__ pos_out, float4* __restrict__ nrm_out,
													 uchar4* __restrict__ rgba_out, DeviceCounters* __restrict__ counters)
{
	// warp = 8x4 pixel tile; CTA = 4x2 tiles = 32x8 pixels
	int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	int const px = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
	int const py = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
	LaneCounters lc = {};
	bool active = px < mp.W && py < mp.H;
	if (active && mp.part_world > 1)
	{
		int const tile = (py / mp.part_th) * mp.part_tiles_x + (px / mp.part_tw);
		active = (tile % mp.part_world) == mp.part_rank;
	}
	uint32_t const index = (uint32_t)py * (uint32_t)mp.W + (uint32_t)px;
	if (active && mp.skip_last_pixel && index == (uint32_t)mp.W * (uint32_t)mp.H - 1u) active = false;   // ThreadPool.cpp:50

	if (active)
	{
		float4 P = make_float4(0.0f, 0.0f, 0.0f, 0.0f), N = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		if (mp.do_march)
		{
			float const z = depth[index];
			if (z != 1.0f)   // `if (z == 1.0f) return;` (RayMarcher.cpp:264)
			{
				lc.covered = 1;
				// pixel CORNER, not centre (RayMarcher.cpp:268-270)
				float const cx = subr(mulr((float)px, mp.two_w_inv), 1.0f);
				float const cy = subr(mulr((float)py, mp.two_h_inv), 1.0f);
				float wh[4];
				mat4_mul_vec4(mp.ipv, cx, cy, z, 1.0f, wh);
				f3 position = divs3(mk3(wh[0], wh[1], wh[2]), wh[3]);
				f3 const cam = mk3(mp.cam[0], mp.cam[1], mp.cam[2]);
				f3 const step = scale3(normalize3(sub3(position, cam)), mp.step_size);
				f3 prev = position;

				for (int i = 0; i < mp.max_steps; i++)
				{
					prev = position;
					position = add3(position, step);

					// empty-space skip (RayMarcher.cpp:282-306); skips do not consume MaxSteps
					int gx, gy, gz;
					bool inside;
					while ((inside = density_cell_of(f, position, gx, gy, gz)) && !density_cell_flag(f, gx, gy, gz))
					{
						// node->Min = m_Min + vec3(x,y,z)*cellWidth; node->Max = Min + vec3(cellWidth) (Dataset.cpp:132-133)
						f3 const nmin = add3(mk3(f.mn.x, f.mn.y, f.mn.z),
											 scale3(mk3((float)gx, (float)gy, (float)gz), f.cell_width));
						f3 const nmax = add3(nmin, mk3(f.cell_width, f.cell_width, f.cell_width));
						prev = position;
						position = add3(intersect_aabb(position, step, nmin, nmax), step);
						lc.skips++;
					}

					if (!inside && mp.early_out)
					{
						// Outside the density grid every particle is farther than h (the grid is the particle
						// AABB padded by h), so density is 0 here.  Each coordinate moves monotonically, so a
						// ray that is outside on an axis and moving away on it can never come back: the
						// remaining samples of the reference are all misses.  Stop; the result is unchanged.
						float const rx = floorf(mulr(subr(position.x, f.mn.x), f.inv_cell_width.x));
						float const ry = floorf(mulr(subr(position.y, f.mn.y), f.inv_cell_width.y));
						float const rz = floorf(mulr(subr(position.z, f.mn.z), f.inv_cell_width.z));
						bool const gone =
							(rx < 0.0f && step.x <= 0.0f) || (rx >= (float)f.gdim.x && step.x >= 0.0f) ||
							(ry < 0.0f && step.y <= 0.0f) || (ry >= (float)f.gdim.y && step.y >= 0.0f) ||
							(rz < 0.0f && step.z <= 0.0f) || (rz >= (float)f.gdim.z && step.z >= 0.0f);
						if (gone) { lc.early_exits = 1; break; }
					}

					f3 grad;
					float const density = eval_density<false>(f, position, grad, lc);
					lc.steps++;

					if (density >= mp.iso)   // RayMarcher.cpp:327
					{
						// optional refinement (north_star item 3; not in the reference): bisect between the last
						// sample below the threshold and the hit sample
						f3 lo = prev, hi = position;
						for (int b = 0; b < mp.bisection_steps; b++)
						{
							f3 const mid = scale3(add3(lo, hi), 0.5f);
							float const dm = eval_density<false>(f, mid, grad, lc);
							lc.steps++;
							if (dm >= mp.iso) hi = mid; else lo = mid;
						}
						position = hi;
						P = make_float4(position.x, position.y, position.z, 1.0f);
						eval_density<true>(f, position, grad, lc);
						f3 const n = normalize3(grad);   // glm::normalize(normal) (RayMarcher.cpp:338)
						N = make_float4(n.x, n.y, n.z, 1.0f);
						lc.hits = 1;
						break;
					}
				}
			}
			pos_out[index] = P;
			nrm_out[index] = N;
		}
		else if (mp.do_shade)
		{
			P = pos_out[index];
			N = nrm_out[index];
		}
		if (mp.do_shade) rgba_out[index] = shade_pixel(mp, px, py, P, N);
	}

	// per-warp counter reduction, one atomic per counter per warp
	uint32_t vals[8] = { lc.covered, lc.hits, lc.steps, lc.skips, lc.candidates, lc.neighbours, lc.early_exits, lc.overflow };
	unsigned long long* dst = reinterpret_cast<unsigned long long*>(counters);
#pragma unroll
	for (int k = 0; k < 8; k++)
	{
		uint32_t const s = __reduce_add_sync(0xffffffffu, vals[k]);
		if (lane == 0 && s) atomicAdd(dst + k, (unsigned long long)s);
	}
}

}  // namespace

int launch_march(Context* ctx, const Frame& f, bool do_march, bool do_shade)
{
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
	mp.part_rank = ctx->part_rank; mp.part_world = ctx->part_world;
	mp.part_tw = ctx->part_tw; mp.part_th = ctx->part_th;
	mp.part_tiles_x = (ctx->width + ctx->part_tw - 1) / ctx->part_tw;
	mp.do_march = do_march ? 1 : 0;
	mp.do_shade = do_shade ? 1 : 0;

	FM_CUDA(cudaMemsetAsync(ctx->d_counters, 0, sizeof(DeviceCounters), ctx->stream));
	dim3 const grid((ctx->width + 31) / 32, (ctx->height + 7) / 8);
	k_march_shade<<<grid, 256, 0, ctx->stream>>>(make_view(f), mp, ctx->d_depth, ctx->d_pos, ctx->d_nrm,
												 ctx->d_rgba_target, ctx->d_counters);
	ctx->kernel_launches += 1;
	FM_CUDA(cudaGetLastError());
	return FR_OK;
}

}  // namespace fm
