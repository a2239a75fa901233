/* fluid_oracle.c -- CPU parity oracle (plain C restatement of the reference's hot path).
 *
 * TEST INFRASTRUCTURE ONLY -- see fluid_oracle.h.  Build: oracle/Makefile
 * (gcc -O2 -ffp-contract=off: every FP32 operation below is one IEEE-754 single operation,
 * written in the order the reference's glm/C++ expressions evaluate them).
 *
 * Citations are relative to /root/reference.
 */
#include "fluid_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------- */
/* minimal pthread parallel-for (the image has no libgomp): dynamic chunks off one counter */

typedef void (*pf_body)(int64_t begin, int64_t end, int tid, void* ctx);

typedef struct
{
	int64_t n, chunk;
	int64_t next;
	pf_body body;
	void* ctx;
} pf_job;

typedef struct { pf_job* job; int tid; } pf_arg;

static void* pf_worker(void* a)
{
	pf_arg* arg = (pf_arg*)a;
	pf_job* j = arg->job;
	for (;;)
	{
		int64_t const b = __atomic_fetch_add(&j->next, j->chunk, __ATOMIC_RELAXED);
		if (b >= j->n) break;
		int64_t const e = b + j->chunk < j->n ? b + j->chunk : j->n;
		j->body(b, e, arg->tid, j->ctx);
	}
	return NULL;
}

static int pf_default_threads(void)
{
	long n = sysconf(_SC_NPROCESSORS_ONLN);
	return n > 0 ? (int)n : 1;
}

static int g_threads = 0;   /* 0 = all online cores */

static void parallel_for(int64_t n, int64_t chunk, pf_body body, void* ctx)
{
	int nt = g_threads > 0 ? g_threads : pf_default_threads();
	if (nt > 256) nt = 256;
	pf_job job = { n, chunk, 0, body, ctx };
	pthread_t th[256];
	pf_arg args[256];
	for (int t = 0; t < nt; t++) { args[t].job = &job; args[t].tid = t; }
	for (int t = 1; t < nt; t++) pthread_create(&th[t], NULL, pf_worker, &args[t]);
	pf_worker(&args[0]);
	for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
}

void fo_set_threads(int threads) { g_threads = threads; }
int fo_get_threads(void) { return g_threads > 0 ? g_threads : pf_default_threads(); }

/* ------------------------------------------------------------------------------------- */
/* small vector helpers (glm semantics: component-wise, one rounding per operation)      */

typedef struct { float x, y, z; } v3;

static inline v3 v3_make(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_mul(v3 a, v3 b) { return v3_make(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3_div(v3 a, v3 b) { return v3_make(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_divs(v3 a, float s) { return v3_make(a.x / s, a.y / s, a.z / s); }
/* glm::dot(vec3): tmp = a*b; tmp.x + tmp.y + tmp.z  (vendor/glm/glm/detail/func_geometric.inl:48-54) */
static inline float v3_dot(v3 a, v3 b) { v3 t = v3_mul(a, b); return (t.x + t.y) + t.z; }
/* glm::min(x,y) = (y < x) ? y : x ; glm::max(x,y) = (x < y) ? y : x  (func_common.inl:17-30) */
static inline float glm_min(float x, float y) { return (y < x) ? y : x; }
static inline float glm_max(float x, float y) { return (x < y) ? y : x; }
/* glm::normalize(v) = v * (1 / sqrt(dot(v,v)))  (func_geometric.inl:82-90, func_exponential.inl:136-139) */
static inline v3 v3_normalize(v3 a) { return v3_scale(a, 1.0f / sqrtf(v3_dot(a, a))); }

/* glm mat4 * vec4, column-major m[c*4+r]: (m0*v0 + m1*v1) + (m2*v2 + m3*v3)
 * (vendor/glm/glm/detail/type_mat4x4.inl:561-572) */
static inline void mat4_mul_vec4(const float* m, const float v[4], float out[4])
{
	for (int r = 0; r < 4; r++)
	{
		float const a0 = m[0 + r] * v[0] + m[4 + r] * v[1];
		float const a1 = m[8 + r] * v[2] + m[12 + r] * v[3];
		out[r] = a0 + a1;
	}
}

int fo_abi_version(void) { return 2; }

/* ------------------------------------------------------------------------------------- */
/* Kernel.cpp                                                                            */

typedef struct { float h, h_squared, h_inv, sig_d; } spline_kernel;

/* CubicSplineKernel::CubicSplineKernel (Kernel.cpp:8-14) */
static spline_kernel spline_make(float h)
{
	spline_kernel k;
	k.h = h;
	k.h_squared = h * h;
	k.h_inv = 1.0f / h;
	k.sig_d = 8.0f / (3.14159265358979323846264338327950288f * h * h * h);
	return k;
}

/* CubicSplineKernel::W (Kernel.cpp:16-32) */
static inline float spline_W(const spline_kernel* k, v3 r)
{
	float q = v3_dot(r, r);
	if (q >= k->h_squared) return 0.0f;
	q = sqrtf(q) * k->h_inv;
	if (q >= 0.5f)
	{
		float const q_ = 1.0f - q;
		return k->sig_d * (2.0f * q_ * q_ * q_);
	}
	return k->sig_d * (6.0f * (q * q * q - q * q) + 1.0f);
}

/* CubicSplineKernel::gradW (Kernel.cpp:34-52); note gradQ = normalize(r) / (|r| * h) as written there */
static inline v3 spline_gradW(const spline_kernel* k, v3 r)
{
	float const rn = v3_dot(r, r);
	if (rn >= k->h_squared) return v3_make(0.0f, 0.0f, 0.0f);
	float const r_length = sqrtf(rn);
	float const q = r_length * k->h_inv;
	v3 const gradQ = v3_divs(v3_normalize(r), r_length * k->h);
	if (q >= 0.5f)
	{
		float const q_ = 1.0f - q;
		return v3_scale(v3_scale(gradQ, -k->sig_d), 6.0f * q_ * q_);
	}
	return v3_scale(v3_scale(gradQ, k->sig_d), 6.0f * (3.0f * q * q - 2.0f * q));
}

float fo_W0(float h) { spline_kernel k = spline_make(h); return k.sig_d; }

float fo_W(float h, const float r[3])
{
	spline_kernel k = spline_make(h);
	return spline_W(&k, v3_make(r[0], r[1], r[2]));
}

void fo_gradW(float h, const float r[3], float out[3])
{
	spline_kernel k = spline_make(h);
	v3 g = spline_gradW(&k, v3_make(r[0], r[1], r[2]));
	out[0] = g.x; out[1] = g.y; out[2] = g.z;
}

/* intersectAABB (RayMarcher.cpp:51-62) */
static inline v3 intersect_aabb(v3 o, v3 d, v3 bmin, v3 bmax)
{
	v3 const tMin = v3_div(v3_sub(bmin, o), d);
	v3 const tMax = v3_div(v3_sub(bmax, o), d);
	v3 const t1 = v3_make(glm_min(tMin.x, tMax.x), glm_min(tMin.y, tMax.y), glm_min(tMin.z, tMax.z));
	v3 const t2 = v3_make(glm_max(tMin.x, tMax.x), glm_max(tMin.y, tMax.y), glm_max(tMin.z, tMax.z));
	float const tNear = glm_max(glm_max(t1.x, t1.y), t1.z);
	float const tFar = glm_min(glm_min(t2.x, t2.y), t2.z);
	float const t = (tNear < tFar) ? tFar : tNear;   /* std::max(tNear, tFar) */
	return t >= 0.0f ? v3_add(o, v3_scale(d, t)) : o;
}

void fo_intersect_aabb(const float o[3], const float d[3], const float bmin[3], const float bmax[3], float out[3])
{
	v3 r = intersect_aabb(v3_make(o[0], o[1], o[2]), v3_make(d[0], d[1], d[2]),
						  v3_make(bmin[0], bmin[1], bmin[2]), v3_make(bmax[0], bmax[1], bmax[2]));
	out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

/* cos(pi/2 * s) for s in [0,1] -- depth.frag:25 `cos(PI_OVER_TWO * sqrt(l2))`.
 * GLSL leaves cos() precision to the implementation (Vulkan: 2^-11 absolute), so the
 * restatement fixes ONE deterministic evaluation, shared op for op with the CUDA kernel:
 * even Taylor polynomial to x^12 in Horner form, |error| < 1e-8 on [0, pi/2]. */
float fo_cos_half_pi(float s)
{
	float const x = 1.57079632679f * s;
	float const x2 = x * x;
	float p = 2.08767569878681e-9f;          /*  1/12! */
	p = p * x2 + -2.75573192239859e-7f;      /* -1/10! */
	p = p * x2 + 2.48015873015873e-5f;       /*  1/8!  */
	p = p * x2 + -1.38888888888889e-3f;      /* -1/6!  */
	p = p * x2 + 4.16666666666667e-2f;       /*  1/4!  */
	p = p * x2 + -0.5f;
	p = p * x2 + 1.0f;
	return p;
}

#include "fluid_oracle_aniso.inc"

/* ------------------------------------------------------------------------------------- */
/* neighbour search: restatement of CompactNSearch as used by Dataset.cpp                */
/* (hash of world-origin cells of size r; see oracle/ref/shim/CompactNSearch.h)          */

typedef struct
{
	float r, r2, inv;
	int32_t kmin[3], kdim[3];    /* dense table over the occupied key range */
	uint32_t* cell_start;        /* kdim product + 1 */
	uint32_t* ids;               /* point ids sorted by cell, ascending id inside a cell */
} nsearch;

static inline void ns_cell_of(const nsearch* s, const float* x, int32_t k[3])
{
	for (int i = 0; i < 3; i++)
	{
		int32_t const t = (int32_t)(s->inv * x[i]);
		k[i] = x[i] >= 0.0f ? t : t - 1;
	}
}

static uint64_t spread21(uint32_t v)
{
	uint64_t x = v & 0x1fffffu;
	x = (x | x << 32) & 0x1f00000000ffffull;
	x = (x | x << 16) & 0x1f0000ff0000ffull;
	x = (x | x << 8) & 0x100f00f00f00f00full;
	x = (x | x << 4) & 0x10c30c30c30c30c3ull;
	x = (x | x << 2) & 0x1249249249249249ull;
	return x;
}

typedef struct { uint64_t code; uint32_t id; } code_id;

static int code_id_cmp(const void* a, const void* b)
{
	const code_id* p = (const code_id*)a;
	const code_id* q = (const code_id*)b;
	if (p->code != q->code) return p->code < q->code ? -1 : 1;
	return p->id < q->id ? -1 : (p->id > q->id ? 1 : 0);
}

/* z_sort + sort_field (Dataset.cpp:58-62): permutes xyz in place into Morton order of the hash cells */
static void ns_z_sort(float r, float* xyz, size_t n)
{
	nsearch tmp;
	memset(&tmp, 0, sizeof tmp);
	tmp.r = r; tmp.r2 = r * r; tmp.inv = 1.0f / r;
	code_id* keys = (code_id*)malloc(n * sizeof(code_id));
	for (size_t i = 0; i < n; i++)
	{
		int32_t k[3];
		ns_cell_of(&tmp, xyz + 3 * i, k);
		uint32_t const bias = (uint32_t)INT_MIN;
		keys[i].code = spread21((uint32_t)k[0] - bias) | spread21((uint32_t)k[1] - bias) << 1 |
			spread21((uint32_t)k[2] - bias) << 2;
		keys[i].id = (uint32_t)i;
	}
	qsort(keys, n, sizeof(code_id), code_id_cmp);   /* total order == stable sort by code */
	float* copy = (float*)malloc(n * 12);
	memcpy(copy, xyz, n * 12);
	for (size_t i = 0; i < n; i++) memcpy(xyz + 3 * i, copy + 3 * (size_t)keys[i].id, 12);
	free(copy);
	free(keys);
}

/* update_point_sets (Dataset.cpp:62): rebuild the cell table from the (permuted) positions */
static void ns_build(nsearch* s, float r, const float* xyz, size_t n)
{
	s->r = r; s->r2 = r * r; s->inv = 1.0f / r;
	int32_t kmax[3] = { INT_MIN, INT_MIN, INT_MIN };
	s->kmin[0] = s->kmin[1] = s->kmin[2] = INT_MAX;
	for (size_t i = 0; i < n; i++)
	{
		int32_t k[3];
		ns_cell_of(s, xyz + 3 * i, k);
		for (int a = 0; a < 3; a++)
		{
			if (k[a] < s->kmin[a]) s->kmin[a] = k[a];
			if (k[a] > kmax[a]) kmax[a] = k[a];
		}
	}
	if (n == 0) { s->kmin[0] = s->kmin[1] = s->kmin[2] = 0; kmax[0] = kmax[1] = kmax[2] = 0; }
	for (int a = 0; a < 3; a++) s->kdim[a] = kmax[a] - s->kmin[a] + 1;
	size_t const ncell = (size_t)s->kdim[0] * (size_t)s->kdim[1] * (size_t)s->kdim[2];
	s->cell_start = (uint32_t*)calloc(ncell + 1, sizeof(uint32_t));
	s->ids = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
	uint32_t* cell_of_point = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
	for (size_t i = 0; i < n; i++)
	{
		int32_t k[3];
		ns_cell_of(s, xyz + 3 * i, k);
		size_t const c = (size_t)(k[0] - s->kmin[0]) +
			(size_t)s->kdim[0] * ((size_t)(k[1] - s->kmin[1]) + (size_t)s->kdim[1] * (size_t)(k[2] - s->kmin[2]));
		cell_of_point[i] = (uint32_t)c;
		s->cell_start[c + 1]++;
	}
	for (size_t c = 0; c < ncell; c++) s->cell_start[c + 1] += s->cell_start[c];
	uint32_t* fill = (uint32_t*)malloc((ncell ? ncell : 1) * sizeof(uint32_t));
	memcpy(fill, s->cell_start, ncell * sizeof(uint32_t));
	for (size_t i = 0; i < n; i++) s->ids[fill[cell_of_point[i]]++] = (uint32_t)i;   /* ascending id per cell */
	free(fill);
	free(cell_of_point);
}

static void ns_free(nsearch* s)
{
	free(s->cell_start);
	free(s->ids);
	s->cell_start = NULL;
	s->ids = NULL;
}

/* NeighborhoodSearch::find_neighbors(point): 27 cells, dj/dk/dl nested, strict d^2 < r^2.
 * Calls visit(id, r_rel) in result order; returns candidates examined. */
typedef struct
{
	uint32_t* out; size_t cap; size_t count; uint64_t candidates;
} ns_result;

static inline void ns_query(const nsearch* s, const float* xyz, const float* x, ns_result* res)
{
	int32_t c[3];
	ns_cell_of(s, x, c);
	for (int dj = -1; dj <= 1; dj++)
	{
		int32_t const k0 = c[0] + dj - s->kmin[0];
		if (k0 < 0 || k0 >= s->kdim[0]) continue;
		for (int dk = -1; dk <= 1; dk++)
		{
			int32_t const k1 = c[1] + dk - s->kmin[1];
			if (k1 < 0 || k1 >= s->kdim[1]) continue;
			for (int dl = -1; dl <= 1; dl++)
			{
				int32_t const k2 = c[2] + dl - s->kmin[2];
				if (k2 < 0 || k2 >= s->kdim[2]) continue;
				size_t const cell = (size_t)k0 + (size_t)s->kdim[0] * ((size_t)k1 + (size_t)s->kdim[1] * (size_t)k2);
				uint32_t const b = s->cell_start[cell], e = s->cell_start[cell + 1];
				res->candidates += e - b;
				for (uint32_t j = b; j < e; j++)
				{
					uint32_t const id = s->ids[j];
					const float* xb = xyz + 3 * (size_t)id;
					float const d0 = x[0] - xb[0], d1 = x[1] - xb[1], d2 = x[2] - xb[2];
					float const l2 = d0 * d0 + d1 * d1 + d2 * d2;
					if (l2 < s->r2)
					{
						if (res->count < res->cap) res->out[res->count] = id;
						res->count++;
					}
				}
			}
		}
	}
}

/* ------------------------------------------------------------------------------------- */
/* Dataset.cpp: Frame                                                                    */

struct fo_frame
{
	size_t n;
	float h, h_ext;
	float* particles;        /* m_Particles after z-sort with r = h */
	float* particles_ext;    /* m_ParticlesExt after z-sort with r = h_ext */
	nsearch search, search_ext;
	v3 mn, mx;               /* m_Min, m_Max */
	/* DensityGrid (Dataset.h:27-35) */
	float cell_width;
	v3 inv_cell_width;
	int32_t gw, gh, gd;
	uint32_t* counts;        /* OctreeNode::NumParticles */
	uint8_t* flags;          /* OctreeNode::Flag */
	spline_kernel kernel;
};

/* Frame::QueryDensityGrid (Dataset.cpp:26-47) */
static inline int64_t frame_query_cell(const fo_frame* f, v3 p)
{
	v3 const rel = v3_mul(v3_sub(p, f->mn), f->inv_cell_width);
	float const fx = floorf(rel.x), fy = floorf(rel.y), fz = floorf(rel.z);
	/* the float range test equals the reference's int32 test for every value an int32 can hold
	 * and is well defined for NaN / huge values, where the reference's cast is not */
	if (fx >= 0.0f && fx < (float)f->gw && fy >= 0.0f && fy < (float)f->gh && fz >= 0.0f && fz < (float)f->gd)
	{
		int32_t const x = (int32_t)fx, y = (int32_t)fy, z = (int32_t)fz;
		return (int64_t)x + (int64_t)y * f->gw + (int64_t)z * f->gw * f->gh;
	}
	return -1;
}

/* Frame::ComputeAABB (Dataset.cpp:78-92) */
static void frame_compute_aabb(fo_frame* f)
{
	v3 mx = v3_make(f->particles[0], f->particles[1], f->particles[2]);
	v3 mn = mx;
	for (size_t i = 0; i < f->n; i++)
	{
		const float* p = f->particles + 3 * i;
		mx = v3_make(glm_max(mx.x, p[0]), glm_max(mx.y, p[1]), glm_max(mx.z, p[2]));
		mn = v3_make(glm_min(mn.x, p[0]), glm_min(mn.y, p[1]), glm_min(mn.z, p[2]));
	}
	float const pad = 1.0f * f->h;
	f->mx = v3_add(mx, v3_make(pad, pad, pad));
	f->mn = v3_sub(mn, v3_make(pad, pad, pad));
}

static int g_count_mode = FO_COUNT_CELL_EXACT;
void fo_set_count_mode(int mode) { g_count_mode = mode; }

/* Frame::BuildDensityGrid (Dataset.cpp:94-165).  NumParticles follows the "cell-exact" reading
 * of the fork-only find_neighbors_box: a particle is counted by the cell that
 * QueryDensityGrid(particle) returns. */
static void frame_build_density_grid(fo_frame* f, float iso_density)
{
	float const cw = 1.0f * f->h;
	f->cell_width = cw;
	f->inv_cell_width = v3_make(1.0f / cw, 1.0f / cw, 1.0f / cw);
	v3 const aabb = v3_sub(f->mx, f->mn);
	int32_t const w = f->gw = (int32_t)ceilf(aabb.x / cw);
	int32_t const h = f->gh = (int32_t)ceilf(aabb.y / cw);
	int32_t const d = f->gd = (int32_t)ceilf(aabb.z / cw);
	size_t const ncell = (size_t)w * (size_t)h * (size_t)d;
	f->counts = (uint32_t*)calloc(ncell ? ncell : 1, sizeof(uint32_t));
	f->flags = (uint8_t*)calloc(ncell ? ncell : 1, 1);

	for (size_t i = 0; i < f->n; i++)
	{
		const float* p = f->particles + 3 * i;
		int64_t const c = frame_query_cell(f, v3_make(p[0], p[1], p[2]));
		if (g_count_mode == FO_COUNT_CELL_EXACT)
		{
			if (c >= 0) f->counts[c]++;
			continue;
		}
		/* FO_COUNT_CENTRE_BOX: what oracle/ref/shim/CompactNSearch.h::find_neighbors_box (mode 0) returns for
		 * the query the reference makes (Dataset.cpp:122-128): centre = m_Min + (vec3(x,y,z) + 0.5) * cellWidth,
		 * particle counted iff centre - r/2 <= p < centre + r/2 on every axis (r = search radius = h), all in
		 * FP32.  Rounding makes these boxes overlap or leave gaps by an ulp, so test the 27 cells around. */
		v3 const rel = v3_mul(v3_sub(v3_make(p[0], p[1], p[2]), f->mn), f->inv_cell_width);
		int32_t const bx = (int32_t)floorf(rel.x), by = (int32_t)floorf(rel.y), bz = (int32_t)floorf(rel.z);
		float const half = 0.5f * f->h;
		for (int32_t x = bx - 1; x <= bx + 1; x++)
			for (int32_t y = by - 1; y <= by + 1; y++)
				for (int32_t z = bz - 1; z <= bz + 1; z++)
				{
					if (x < 0 || x >= w || y < 0 || y >= h || z < 0 || z >= d) continue;
					v3 const centre = v3_add(f->mn, v3_scale(v3_add(v3_make((float)x, (float)y, (float)z), v3_make(0.5f, 0.5f, 0.5f)), cw));
					if (p[0] >= centre.x - half && p[0] < centre.x + half && p[1] >= centre.y - half && p[1] < centre.y + half &&
						p[2] >= centre.z - half && p[2] < centre.z + half)
						f->counts[(size_t)x + (size_t)y * w + (size_t)z * w * h]++;
				}
	}

	float const W0 = f->kernel.sig_d;
	for (int32_t z = 0; z < d; z++)
		for (int32_t y = 0; y < h; y++)
			for (int32_t x = 0; x < w; x++)
			{
				float N_c = 0.0f;
				for (int32_t dx = -1; dx <= 1; dx++)
					for (int32_t dy = -1; dy <= 1; dy++)
						for (int32_t dz = -1; dz <= 1; dz++)
						{
							int32_t const xx = x + dx, yy = y + dy, zz = z + dz;
							if (xx < 0 || xx >= w || yy < 0 || yy >= h || zz < 0 || zz >= d) continue;
							float const C = 1000.0f;
							float const r = (float)(abs(dx) + abs(dy) + abs(dz));
							float const weight = expf(-C * r);
							N_c += weight * (float)f->counts[(size_t)xx + (size_t)yy * w + (size_t)zz * w * h];
						}
				float const rho = N_c * W0;
				f->flags[(size_t)x + (size_t)y * w + (size_t)z * w * h] = rho > iso_density;
			}
}

fo_frame* fo_frame_create(const float* xyz, size_t n, float h, float h_mult)
{
	if (n == 0 || !xyz || !(h > 0.0f)) return NULL;
	fo_frame* f = (fo_frame*)calloc(1, sizeof(fo_frame));
	f->n = n;
	f->h = h;
	f->h_ext = h_mult * h;
	f->kernel = spline_make(h);
	f->particles = (float*)malloc(n * 12);
	f->particles_ext = (float*)malloc(n * 12);
	memcpy(f->particles, xyz, n * 12);
	memcpy(f->particles_ext, xyz, n * 12);
	/* Frame::BuildSearch (Dataset.cpp:49-76) */
	ns_z_sort(f->h, f->particles, n);
	ns_build(&f->search, f->h, f->particles, n);
	ns_z_sort(f->h_ext, f->particles_ext, n);
	ns_build(&f->search_ext, f->h_ext, f->particles_ext, n);
	frame_compute_aabb(f);
	frame_build_density_grid(f, 1.0f);   /* BuildDensityGrid(1), Dataset.cpp:23 */
	return f;
}

void fo_frame_destroy(fo_frame* f)
{
	if (!f) return;
	ns_free(&f->search);
	ns_free(&f->search_ext);
	free(f->particles);
	free(f->particles_ext);
	free(f->counts);
	free(f->flags);
	free(f);
}

size_t fo_frame_num_particles(const fo_frame* f) { return f->n; }

void fo_frame_info(const fo_frame* f, float mn[3], float mx[3], int32_t dims[3])
{
	mn[0] = f->mn.x; mn[1] = f->mn.y; mn[2] = f->mn.z;
	mx[0] = f->mx.x; mx[1] = f->mx.y; mx[2] = f->mx.z;
	dims[0] = f->gw; dims[1] = f->gh; dims[2] = f->gd;
}

void fo_frame_particles(const fo_frame* f, float* xyz) { memcpy(xyz, f->particles, f->n * 12); }
void fo_frame_particles_ext(const fo_frame* f, float* xyz) { memcpy(xyz, f->particles_ext, f->n * 12); }

void fo_frame_grid(const fo_frame* f, uint32_t* counts, uint8_t* flags)
{
	size_t const ncell = (size_t)f->gw * (size_t)f->gh * (size_t)f->gd;
	if (counts) memcpy(counts, f->counts, ncell * sizeof(uint32_t));
	if (flags) memcpy(flags, f->flags, ncell);
}

int64_t fo_query_cell(const fo_frame* f, const float p[3]) { return frame_query_cell(f, v3_make(p[0], p[1], p[2])); }

size_t fo_neighbors(const fo_frame* f, const float p[3], int ext, uint32_t* out, size_t cap)
{
	ns_result res = { out, cap, 0, 0 };
	if (ext) ns_query(&f->search_ext, f->particles_ext, p, &res);
	else ns_query(&f->search, f->particles, p, &res);
	return res.count;
}

/* ------------------------------------------------------------------------------------- */
/* depth pre-pass                                                                        */
/* CollectRenderData (AdvancedRenderer.cpp:447-485): per particle a quad p +- h*System[0] +- h*System[1],
 * UV in [-1,1]^2.  System = transpose(View) (CameraController3D.cpp:80), so the quad lies in the
 * plane view-z = z_c and spans +-h in view x/y around the particle's view position.
 * depth.vert:20-27: ViewPosition = (View*p).xyz/w.  depth.frag:19-33: l2 = dot(uv,uv), discard if
 * l2 > 1, off = cos(pi/2*sqrt(l2)), pView = ViewPosition - Radius*(0,0,off), depth = (P*pView).z/w.
 * DepthRenderPass.cpp:54 clear 1.0, :155 cull none, :177-179 depth test Less.
 * The rasteriser evaluates fragments at pixel centres (px+.5, py+.5); on the quad plane the
 * pixel's view ray has x_v = ndc_x*z_c/P00, y_v = ndc_y*z_c/P11, hence uv = (x_v-x_c, y_v-y_c)/h.
 * This analytic form IS the restatement (a rasteriser's interpolation is not bit-defined).
 * Quads entirely outside the near/far range are clipped; gl_FragDepth is clamped to [0,1]. */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

typedef struct
{
	const fo_frame* f; int32_t W, H; const float* view;
	float P00, P11, P22, P32;
	uint32_t* dbits;
} depth_ctx;

static void depth_body(int64_t begin, int64_t end, int tid, void* vctx)
{
	(void)tid;
	const depth_ctx* c = (const depth_ctx*)vctx;
	const fo_frame* f = c->f;
	int32_t const W = c->W, H = c->H;
	const float* view = c->view;
	float const P00 = c->P00, P11 = c->P11, P22 = c->P22, P32 = c->P32;
	float const h = f->h, h_inv = 1.0f / h;
	float const two_w_inv = 2.0f / (float)W, two_h_inv = 2.0f / (float)H;
	float const half_w = 0.5f * (float)W, half_h = 0.5f * (float)H;
	uint32_t* dbits = c->dbits;
	for (int64_t i = begin; i < end; i++)
	{
		const float* p = f->particles + 3 * (size_t)i;
		float const pv[4] = { p[0], p[1], p[2], 1.0f };
		float vc[4];
		mat4_mul_vec4(view, pv, vc);
		float const x_c = vc[0] / vc[3], y_c = vc[1] / vc[3], z_c = vc[2] / vc[3];
		if (!(z_c > 0.0f)) continue;
		float const quad_depth = (P22 * z_c + P32) / z_c;
		if (!(quad_depth >= 0.0f && quad_depth <= 1.0f)) continue;   /* clipped by near / far */

		/* conservative pixel bounding box of the disc (any superset gives the same image) */
		float const cx = (P00 * x_c / z_c + 1.0f) * half_w;
		float const cy = (P11 * y_c / z_c + 1.0f) * half_h;
		float const rx = fabsf(P00) * h / z_c * half_w + 1.0f;
		float const ry = fabsf(P11) * h / z_c * half_h + 1.0f;
		float const fx0 = floorf(cx - rx - 0.5f), fx1 = ceilf(cx + rx - 0.5f);
		float const fy0 = floorf(cy - ry - 0.5f), fy1 = ceilf(cy + ry - 0.5f);
		if (!(fx1 >= 0.0f && fy1 >= 0.0f && fx0 <= (float)(W - 1) && fy0 <= (float)(H - 1))) continue;
		int32_t const x0 = fx0 < 0.0f ? 0 : (int32_t)fx0;
		int32_t const y0 = fy0 < 0.0f ? 0 : (int32_t)fy0;
		int32_t const x1 = fx1 > (float)(W - 1) ? W - 1 : (int32_t)fx1;
		int32_t const y1 = fy1 > (float)(H - 1) ? H - 1 : (int32_t)fy1;

		/* uv = ((x_v - x_c)/h, (y_v - y_c)/h) with x_v = ndc_x*z_c/P00 folded into one
		 * multiply-subtract per pixel: u = ndc_x*ax - bx, ax = z_c/(P00*h), bx = x_c/h */
		float const ax = z_c / (P00 * h), bx = x_c * h_inv;
		float const ay = z_c / (P11 * h), by = y_c * h_inv;
		for (int32_t py = y0; py <= y1; py++)
		{
			float const ndc_y = ((float)py + 0.5f) * two_h_inv - 1.0f;
			float const v = ndc_y * ay - by;
			float const vv = v * v;
			for (int32_t px = x0; px <= x1; px++)
			{
				float const ndc_x = ((float)px + 0.5f) * two_w_inv - 1.0f;
				float const u = ndc_x * ax - bx;
				float const l2 = u * u + vv;
				if (l2 > 1.0f) continue;   /* `if (l2 > 1) discard;`  (NaN keeps, as in GLSL) */
				float const off = fo_cos_half_pi(sqrtf(l2));
				float const zf = z_c - h * off;
				float d = (P22 * zf + P32) / zf;
				d = d < 0.0f ? 0.0f : (d > 1.0f ? 1.0f : d);
				if (!(d < 1.0f)) continue;   /* compare Less against the clear value */
				/* depth in [0,1): IEEE order == unsigned order */
				uint32_t const nb = f2u(d);
				uint32_t* cell = dbits + (size_t)py * (size_t)W + (size_t)px;
				uint32_t cur = __atomic_load_n(cell, __ATOMIC_RELAXED);
				while (nb < cur &&
					   !__atomic_compare_exchange_n(cell, &cur, nb, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
			}
		}
	}
}

int fo_depth_prepass(const fo_frame* f, int32_t W, int32_t H, const float view[16], const float proj[16], float* depth)
{
	if (!f || W <= 0 || H <= 0) return -1;
	/* only the perspective structure glm::perspectiveLH_ZO scaled by (1,-1,1) produces is supported
	 * (Camera3D.cpp:11, vendor/glm/glm/ext/matrix_clip_space.inl:265-278) */
	static const int zero_idx[] = { 1, 2, 3, 4, 6, 7, 8, 9, 12, 13, 15 };
	for (size_t i = 0; i < sizeof zero_idx / sizeof zero_idx[0]; i++)
		if (proj[zero_idx[i]] != 0.0f) return -2;
	if (proj[11] != 1.0f) return -2;
	float const P00 = proj[0], P11 = proj[5], P22 = proj[10], P32 = proj[14];
	size_t const npix = (size_t)W * (size_t)H;
	uint32_t* dbits = (uint32_t*)depth;
	for (size_t i = 0; i < npix; i++) depth[i] = 1.0f;

	depth_ctx ctx = { f, W, H, view, P00, P11, P22, P32, dbits };
	parallel_for((int64_t)f->n, 1024, depth_body, &ctx);
	return 0;
}

/* ------------------------------------------------------------------------------------- */
/* the march: RayMarcher::PerPixel_Isotropic (RayMarcher.cpp:256-344)                    */

#define FO_MAX_NEIGHBORS 8192   /* RayMarcher.cpp:14 */

typedef struct
{
	uint32_t ids[FO_MAX_NEIGHBORS];
	v3 rel[FO_MAX_NEIGHBORS];    /* ThreadLocals::NeighborPositions_Rel (RayMarcher.cpp:16-23) */
	v3 abs_ext[FO_MAX_NEIGHBORS]; /* ThreadLocals::NeighborPositions_AbsExt */
	float weights[FO_MAX_NEIGHBORS];
} march_locals;

static void march_pixel_isotropic(const fo_frame* f, int32_t W, const fo_settings* s,
								  float two_w_inv, float two_h_inv, const float* ipv, v3 cam,
								  const float* depth, float* pos4, float* nrm4, float* band, uint32_t* steps,
								  uint32_t index, march_locals* loc, fo_counters* c)
{
	float const z = depth[index];
	memset(pos4 + 4 * (size_t)index, 0, 16);
	memset(nrm4 + 4 * (size_t)index, 0, 16);
	if (band) band[index] = INFINITY;
	if (steps) steps[index] = 0;
	if (z == 1.0f) return;
	c->covered_rays++;

	/* pixel CORNER, not centre (RayMarcher.cpp:268-270) */
	float const clip[4] = { (float)(index % (uint32_t)W) * two_w_inv - 1.0f,
							(float)(index / (uint32_t)W) * two_h_inv - 1.0f, z, 1.0f };
	float wh[4];
	mat4_mul_vec4(ipv, clip, wh);
	v3 position = v3_divs(v3_make(wh[0], wh[1], wh[2]), wh[3]);
	v3 const step = v3_scale(v3_normalize(v3_sub(position, cam)), s->step_size);
	float band_min = INFINITY;
	uint32_t nsteps = 0;

	for (int i = 0; i < s->max_steps; i++)
	{
		position = v3_add(position, step);

		/* empty-space skip (RayMarcher.cpp:282-306) */
		int64_t cell;
		while ((cell = frame_query_cell(f, position)) >= 0 && !f->flags[cell])
		{
			int32_t const cx = (int32_t)(cell % f->gw);
			int32_t const cy = (int32_t)((cell / f->gw) % f->gh);
			int32_t const cz = (int32_t)(cell / ((int64_t)f->gw * f->gh));
			/* node->Min = m_Min + vec3(x,y,z)*cellWidth; node->Max = Min + vec3(cellWidth) (Dataset.cpp:132-133) */
			v3 const nmin = v3_add(f->mn, v3_scale(v3_make((float)cx, (float)cy, (float)cz), f->cell_width));
			v3 const nmax = v3_add(nmin, v3_make(f->cell_width, f->cell_width, f->cell_width));
			position = v3_add(intersect_aabb(position, step, nmin, nmax), step);
			c->skip_iterations++;
		}

		/* Dataset::GetNeighbors (Dataset.cpp:272-280) */
		float const q[3] = { position.x, position.y, position.z };
		ns_result res = { loc->ids, FO_MAX_NEIGHBORS, 0, 0 };
		ns_query(&f->search, f->particles, q, &res);
		uint32_t const nn = res.count < FO_MAX_NEIGHBORS ? (uint32_t)res.count : FO_MAX_NEIGHBORS;
		c->candidates += res.candidates;
		c->neighbours += nn;
		c->ray_steps++;
		if (cell < 0) c->steps_outside_grid++;
		nsteps++;

		for (uint32_t k = 0; k < nn; k++)
		{
			const float* xb = f->particles + 3 * (size_t)loc->ids[k];
			loc->rel[k] = v3_sub(v3_make(xb[0], xb[1], xb[2]), position);
		}

		float density = 0.0f;
		for (uint32_t k = 0; k < nn; k++) density += spline_W(&f->kernel, loc->rel[k]);

		float const dist = fabsf(density - s->iso_density);
		if (dist < band_min) band_min = dist;

		if (density >= s->iso_density)
		{
			float* P = pos4 + 4 * (size_t)index;
			P[0] = position.x; P[1] = position.y; P[2] = position.z; P[3] = 1.0f;
			v3 normal = v3_make(0.0f, 0.0f, 0.0f);
			for (uint32_t k = 0; k < nn; k++) normal = v3_add(normal, spline_gradW(&f->kernel, loc->rel[k]));
			normal = v3_normalize(normal);
			float* N = nrm4 + 4 * (size_t)index;
			N[0] = normal.x; N[1] = normal.y; N[2] = normal.z; N[3] = 1.0f;
			c->hit_rays++;
			break;
		}
	}
	if (band) band[index] = band_min;
	if (steps) steps[index] = nsteps;
}

/* RayMarcher::PerPixel_Anisotropic (RayMarcher.cpp:346-423) */
static void march_pixel_anisotropic(const fo_frame* f, int32_t W, const fo_settings* s,
									float two_w_inv, float two_h_inv, const float* ipv, v3 cam,
									const float* depth, float* pos4, float* nrm4, float* band, uint32_t* steps,
									uint32_t index, march_locals* loc, fo_counters* c)
{
	float const z = depth[index];
	memset(pos4 + 4 * (size_t)index, 0, 16);
	memset(nrm4 + 4 * (size_t)index, 0, 16);
	if (band) band[index] = INFINITY;
	if (steps) steps[index] = 0;
	if (z == 1.0f) return;
	c->covered_rays++;

	float const clip[4] = { (float)(index % (uint32_t)W) * two_w_inv - 1.0f,
							(float)(index / (uint32_t)W) * two_h_inv - 1.0f, z, 1.0f };
	float wh[4];
	mat4_mul_vec4(ipv, clip, wh);
	v3 position = v3_divs(v3_make(wh[0], wh[1], wh[2]), wh[3]);
	v3 const step = v3_scale(v3_normalize(v3_sub(position, cam)), s->step_size);
	float band_min = INFINITY;
	uint32_t nsteps = 0;
	aniso_kernel const ak = aniso_make(f->h);            /* Dataset::m_AnisotropicKernel(ParticleRadius) */
	float const h2 = f->h * f->h;                        /* ParticleRadius * ParticleRadius (RayMarcher.cpp:393) */
	float const particle_radius_inv = 1.0f / f->h;       /* Dataset::ParticleRadiusInv */

	for (int i = 0; i < s->max_steps; i++)
	{
		position = v3_add(position, step);

		int64_t cell;
		while ((cell = frame_query_cell(f, position)) >= 0 && !f->flags[cell])
		{
			int32_t const cx = (int32_t)(cell % f->gw);
			int32_t const cy = (int32_t)((cell / f->gw) % f->gh);
			int32_t const cz = (int32_t)(cell / ((int64_t)f->gw * f->gh));
			v3 const nmin = v3_add(f->mn, v3_scale(v3_make((float)cx, (float)cy, (float)cz), f->cell_width));
			v3 const nmax = v3_add(nmin, v3_make(f->cell_width, f->cell_width, f->cell_width));
			position = v3_add(intersect_aabb(position, step, nmin, nmax), step);
			c->skip_iterations++;
		}

		/* Dataset::GetNeighborsExt (Dataset.cpp:282-290): the r = h_ext search over m_ParticlesExt */
		float const q[3] = { position.x, position.y, position.z };
		ns_result res = { loc->ids, FO_MAX_NEIGHBORS, 0, 0 };
		ns_query(&f->search_ext, f->particles_ext, q, &res);
		uint32_t const n_ext = res.count < FO_MAX_NEIGHBORS ? (uint32_t)res.count : FO_MAX_NEIGHBORS;
		uint32_t nn = 0;
		for (uint32_t k = 0; k < n_ext; k++)
		{
			const float* xb = f->particles_ext + 3 * (size_t)loc->ids[k];
			loc->abs_ext[k] = v3_make(xb[0], xb[1], xb[2]);
			v3 const r = v3_sub(loc->abs_ext[k], position);
			if (v3_dot(r, r) < h2) loc->rel[nn++] = r;
		}
		c->candidates += res.candidates;
		c->neighbours += nn;
		c->ray_steps++;
		if (cell < 0) c->steps_outside_grid++;
		nsteps++;

		float G[3][3];
		an_wpca(f->h_ext, particle_radius_inv, s, position, loc->abs_ext, n_ext, loc->weights, G);
		float const detG = an_det3(G);
		float density = 0.0f;
		for (uint32_t k = 0; k < nn; k++) density += aniso_W(&ak, G, detG, loc->rel[k]);

		float const dist = fabsf(density - s->iso_density);
		if (dist < band_min) band_min = dist;

		if (density >= s->iso_density)
		{
			float* P = pos4 + 4 * (size_t)index;
			P[0] = position.x; P[1] = position.y; P[2] = position.z; P[3] = 1.0f;
			v3 normal = v3_make(0.0f, 0.0f, 0.0f);
			for (uint32_t k = 0; k < nn; k++) normal = v3_add(normal, aniso_gradW(&ak, G, detG, loc->rel[k]));
			normal = v3_normalize(normal);
			float* N = nrm4 + 4 * (size_t)index;
			N[0] = normal.x; N[1] = normal.y; N[2] = normal.z; N[3] = 1.0f;
			c->hit_rays++;
			break;
		}
	}
	if (band) band[index] = band_min;
	if (steps) steps[index] = nsteps;
}

typedef struct
{
	const fo_frame* f; int32_t W; const fo_settings* s;
	float two_w_inv, two_h_inv; const float* ipv; v3 cam;
	const float* depth; float* pos4; float* nrm4; float* band; uint32_t* steps;
	fo_counters per_thread[256];
} march_ctx;

static void march_body(int64_t begin, int64_t end, int tid, void* vctx)
{
	march_ctx* c = (march_ctx*)vctx;
	march_locals* loc = (march_locals*)malloc(sizeof(march_locals));
	for (int64_t i = begin; i < end; i++)
	{
		if (c->s->anisotropic)
			march_pixel_anisotropic(c->f, c->W, c->s, c->two_w_inv, c->two_h_inv, c->ipv, c->cam, c->depth,
									c->pos4, c->nrm4, c->band, c->steps, (uint32_t)i, loc, &c->per_thread[tid]);
		else
			march_pixel_isotropic(c->f, c->W, c->s, c->two_w_inv, c->two_h_inv, c->ipv, c->cam, c->depth,
								  c->pos4, c->nrm4, c->band, c->steps, (uint32_t)i, loc, &c->per_thread[tid]);
	}
	free(loc);
}

int fo_march(const fo_frame* f, int32_t W, int32_t H, const fo_settings* s,
			 const float inv_proj_view[16], const float cam_pos[3], const float* depth,
			 float* pos4, float* nrm4, float* band, uint32_t* steps, fo_counters* counters,
			 int threads)
{
	if (!f || !s || W <= 0 || H <= 0) return -1;
	march_ctx* c = (march_ctx*)calloc(1, sizeof(march_ctx));
	c->f = f; c->W = W; c->s = s;
	c->two_w_inv = 2.0f / (float)W;   /* RayMarcher.cpp:88-89 */
	c->two_h_inv = 2.0f / (float)H;
	c->ipv = inv_proj_view;
	c->cam = v3_make(cam_pos[0], cam_pos[1], cam_pos[2]);
	c->depth = depth; c->pos4 = pos4; c->nrm4 = nrm4; c->band = band; c->steps = steps;
	int64_t const total = (int64_t)W * (int64_t)H;
	int const saved = g_threads;
	if (threads > 0) g_threads = threads;
	parallel_for(total, 256, march_body, c);
	g_threads = saved;
	fo_counters sum;
	memset(&sum, 0, sizeof sum);
	sum.pixels = (uint64_t)total;
	for (int t = 0; t < 256; t++)
	{
		const fo_counters* p = &c->per_thread[t];
		sum.covered_rays += p->covered_rays; sum.hit_rays += p->hit_rays; sum.ray_steps += p->ray_steps;
		sum.skip_iterations += p->skip_iterations; sum.candidates += p->candidates; sum.neighbours += p->neighbours;
		sum.steps_outside_grid += p->steps_outside_grid;
	}
	free(c);
	if (counters) *counters = sum;
	return 0;
}

/* ------------------------------------------------------------------------------------- */
/* shading: composition.frag                                                             */

/* sampleFloor (composition.frag:37-46) */
static inline void sample_floor(v3 a, v3 r, float out[4])
{
	float const FLOOR_HEIGHT = -1.0f;
	/* b = a + r * (FLOOR_HEIGHT - a.y) / r.y : (r * s) / r.y, left to right */
	float const sN = FLOOR_HEIGHT - a.y;
	v3 const b = v3_add(a, v3_divs(v3_scale(r, sN), r.y));
	/* mod(x, 2) = x - 2*floor(x/2); step(edge = m, x = 1) = x < edge ? 0 : 1 */
	float const mx = b.x - 2.0f * floorf(b.x / 2.0f);
	float const mz = b.z - 2.0f * floorf(b.z / 2.0f);
	float const fx = (1.0f < mx) ? 0.0f : 1.0f;
	float const fy = (1.0f < mz) ? 0.0f : 1.0f;
	float const g = 0.25f + (fx + fy) / 4.0f;
	out[0] = g; out[1] = g; out[2] = g; out[3] = 0.5f;
}

static inline uint8_t unorm8(float x)
{
	if (!(x > 0.0f)) return 0;      /* also NaN */
	if (x >= 1.0f) return 255;
	return (uint8_t)(x * 255.0f + 0.5f);
}

/* linear -> sRGB transfer a *_SRGB colour attachment applies on write (RendererInit2.cpp:50) */
static inline float srgb_encode(float c)
{
	if (!(c > 0.0f)) return 0.0f;
	if (c >= 1.0f) return 1.0f;
	return c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
}

typedef struct
{
	int32_t W, H; const float* pos4; const float* nrm4; const float* ipv; v3 cam, dir;
	float* color4; uint8_t* rgba8;
} shade_ctx;

static void shade_body(int64_t begin, int64_t end, int tid, void* vctx)
{
	(void)tid;
	const shade_ctx* c = (const shade_ctx*)vctx;
	int32_t const W = c->W, H = c->H;
	const float* pos4 = c->pos4; const float* nrm4 = c->nrm4; const float* inv_proj_view = c->ipv;
	v3 const cam = c->cam, dir = c->dir;
	float* color4 = c->color4; uint8_t* rgba8 = c->rgba8;
	float const diffuse[4] = { 120.0f / 255.0f, 185.0f / 255.0f, 255.0f / 255.0f, 255.0f / 255.0f };   /* composition.frag:3 */
	for (int32_t py = (int32_t)begin; py < (int32_t)end; py++)
		for (int32_t px = 0; px < W; px++)
		{
			size_t const idx = (size_t)py * (size_t)W + (size_t)px;
			/* fullscreen.vert: UV interpolates to the pixel centre */
			float const u = ((float)px + 0.5f) / (float)W;
			float const v = ((float)py + 0.5f) / (float)H;
			/* viewRay() (composition.frag:59-66): a far-plane POINT used as a direction */
			float const clip[4] = { 2.0f * u - 1.0f, 2.0f * v - 1.0f, 1.0f, 1.0f };
			float wh[4];
			mat4_mul_vec4(inv_proj_view, clip, wh);
			v3 const view_ray = v3_divs(v3_make(wh[0], wh[1], wh[2]), wh[3]);
			float color[4];
			const float* P = pos4 + 4 * idx;
			if (P[3] == 0.0f)
			{
				float fl[4];
				sample_floor(cam, view_ray, fl);
				for (int k = 0; k < 4; k++) color[k] = 0.75f * fl[k];
			}
			else
			{
				v3 const world = v3_make(P[0], P[1], P[2]);
				const float* Nn = nrm4 + 4 * idx;
				v3 const normal = v3_make(Nn[0], Nn[1], Nn[2]);
				/* refract(I, N, eta) (GLSL): k = 1 - eta^2 (1 - dot(N,I)^2); k < 0 ? 0 : eta*I - (eta*dot(N,I) + sqrt(k))*N */
				v3 const I = v3_normalize(view_ray);
				float const eta = 1.333f;
				float const dotNI = v3_dot(normal, I);
				float const k = 1.0f - eta * eta * (1.0f - dotNI * dotNI);
				v3 refracted = v3_make(0.0f, 0.0f, 0.0f);
				if (k >= 0.0f) refracted = v3_sub(v3_scale(I, eta), v3_scale(normal, eta * dotNI + sqrtf(k)));
				float fl[4];
				sample_floor(world, refracted, fl);
				float const fv = -v3_dot(dir, normal);
				float const amb = 0.15f + 1.0f - fv * fv;
				for (int kk = 0; kk < 4; kk++) color[kk] = fv * fl[kk] + amb * diffuse[kk];
			}
			if (color4) memcpy(color4 + 4 * idx, color, 16);
			if (rgba8)
			{
				rgba8[4 * idx + 0] = unorm8(srgb_encode(color[0]));
				rgba8[4 * idx + 1] = unorm8(srgb_encode(color[1]));
				rgba8[4 * idx + 2] = unorm8(srgb_encode(color[2]));
				rgba8[4 * idx + 3] = unorm8(color[3]);
			}
		}
}

void fo_shade(int32_t W, int32_t H, const float* pos4, const float* nrm4,
			  const float inv_proj_view[16], const float cam_pos[3], const float cam_dir[3],
			  float* color4, uint8_t* rgba8)
{
	shade_ctx c = { W, H, pos4, nrm4, inv_proj_view, v3_make(cam_pos[0], cam_pos[1], cam_pos[2]),
					v3_make(cam_dir[0], cam_dir[1], cam_dir[2]), color4, rgba8 };
	parallel_for(H, 8, shade_body, &c);
}

/* ------------------------------------------------------------------------------------- */
/* screen-space smoothing (f5): GaussRenderPass.cpp:15-66, gauss.frag:28-47,             */
/* composition.frag:50-57,87-104, fullscreen.vert:30-42                                  */

void fo_gauss_kernel(int32_t n, float* out)
{
	float const E = (float)2.7182818284590452353602874713527;      /* powf(E, ...) with E a double macro: float args */
	float sum = 0.0f;
	for (int32_t i = 0; i <= n; i++)
		for (int32_t j = 0; j <= n; j++)
		{
			float const x = sqrtf((float)i * (float)i + (float)j * (float)j);
			float const y = powf(E, -0.5f * x * x);
			out[j + i * (n + 1)] = y;
			if (i == 0 && j == 0) sum += y;
			else if (i != 0 && j != 0) sum += 4.0f * y;
			else sum += 2.0f * y;
		}
	for (int32_t i = 0; i <= n; i++)
		for (int32_t j = 0; j <= n; j++) out[j + i * (n + 1)] /= sum;
}

static inline int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

void fo_gauss_depth(int32_t W, int32_t H, const float* depth, int32_t n, float* smoothed)
{
	float* k = (float*)malloc(sizeof(float) * (size_t)(n + 1) * (size_t)(n + 1));
	fo_gauss_kernel(n, k);
	for (int32_t y = 0; y < H; y++)
		for (int32_t x = 0; x < W; x++)
		{
			float sum = 0.0f;
			for (int32_t i = -n; i <= n; i++)
				for (int32_t j = -n; j <= n; j++)
				{
					int32_t const index = abs(i) * (n + 1) + abs(j);
					float const t = k[index] * depth[(size_t)clampi(y + j, 0, H - 1) * W + clampi(x + i, 0, W - 1)];
					sum = sum + t;
				}
			smoothed[(size_t)y * W + x] = sum;
		}
	free(k);
}

static v3 smoothed_position(int32_t W, int32_t H, const float* smoothed, const float* inv_proj, float u, float v)
{
	/* the texel a linear sampler returns at a texel centre: floor(uv * extent), clamp-to-edge */
	int32_t const tx = clampi((int32_t)floorf(u * (float)W), 0, W - 1), ty = clampi((int32_t)floorf(v * (float)H), 0, H - 1);
	float const clip[4] = { 2.0f * u - 1.0f, 2.0f * v - 1.0f, smoothed[(size_t)ty * W + tx], 1.0f };
	float h[4];
	mat4_mul_vec4(inv_proj, clip, h);
	return v3_divs(v3_make(h[0], h[1], h[2]), h[3]);
}

void fo_sobel_normals(int32_t W, int32_t H, const float* smoothed, const float inv_proj[16], float* nrm4)
{
	float const tw = 1.0f / (float)W, th = 1.0f / (float)H;
	for (int32_t y = 0; y < H; y++)
		for (int32_t x = 0; x < W; x++)
		{
			float const u = ((float)x + 0.5f) / (float)W, v = ((float)y + 0.5f) / (float)H;
			v3 const tl = smoothed_position(W, H, smoothed, inv_proj, u - tw, v - th), tm = smoothed_position(W, H, smoothed, inv_proj, u, v - th),
				tr = smoothed_position(W, H, smoothed, inv_proj, u + tw, v - th), ml = smoothed_position(W, H, smoothed, inv_proj, u - tw, v),
				mr = smoothed_position(W, H, smoothed, inv_proj, u + tw, v), bl = smoothed_position(W, H, smoothed, inv_proj, u - tw, v + th),
				bm = smoothed_position(W, H, smoothed, inv_proj, u, v + th), br = smoothed_position(W, H, smoothed, inv_proj, u + tw, v + th);
			v3 const dx = v3_sub(v3_sub(v3_sub(v3_add(v3_add(v3_scale(tr, 1.0f), v3_scale(mr, 2.0f)), v3_scale(br, 1.0f)), v3_scale(tl, 1.0f)),
										v3_scale(ml, 2.0f)), v3_scale(bl, 1.0f));
			v3 const dy = v3_sub(v3_sub(v3_sub(v3_add(v3_add(v3_scale(bl, 1.0f), v3_scale(bm, 2.0f)), v3_scale(br, 1.0f)), v3_scale(tl, 1.0f)),
										v3_scale(tm, 2.0f)), v3_scale(tr, 1.0f));
			v3 const c = v3_make(dx.y * dy.z - dy.y * dx.z, dx.z * dy.x - dy.z * dx.x, dx.x * dy.y - dy.x * dx.y);      /* glm / GLSL cross */
			v3 const nrm = v3_normalize(c);
			float* o = nrm4 + 4 * ((size_t)y * W + x);
			o[0] = nrm.x; o[1] = nrm.y; o[2] = nrm.z; o[3] = 1.0f;
		}
}
