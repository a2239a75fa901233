// fm_bgeo.cu -- particle files of a sequence: classic Houdini .bgeo (version 5), the format the reference's
// datasets come in (SPlisHSPlasH exports, read through partio).
//
// Replaces, for this path, Dataset::Dataset's file loop + Partio::read + Dataset::ReadFile
// (src/app/Dataset.cpp:169-227, 292-306: files <prefix><i><suffix>, i = 1, 2, ... while they exist; of each file only
// the "position" attribute is used) and restates the reader/writer of vendor/partio/src/io/BGEO.cpp:200-290, 304-460
// (gzip detection: vendor/partio/src/io/ZIP.cpp).  Layout, everything big-endian:
//
//   "Bgeo" 'V' i32 version(5) i32 nPoints nPrims nPointGroups nPrimGroups nPointAttrib nVertexAttrib nPrimAttrib nAttrib
//   nPointAttrib x { u16 len, name, u16 size, i32 type; type 0 float / 1 int / 5 vector: size x i32 defaults;
//                    type 4 indexed string: i32 count, count x { u16 len, text } }
//   nPoints x { f32 x, y, z, w, then the attributes' words }            <- the only block this path needs
//   primitives, detail attributes, 0x00 0xff
//
// Once rendering costs a fraction of a millisecond, decoding the file is the expensive part of a frame.
// So the host only finds the point block (header + attribute table, a few hundred bytes) and copies the block
// verbatim to the GPU; the big-endian -> little-endian swap and the stride-(4 + attributes) -> packed xyz gather are
// a kernel (k_bgeo_unpack), straight into the frame build.  fr_bgeo_read is the same decode for host consumers.
#include "fm_internal.h"

#include <errno.h>
#include <stdlib.h>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <string>
#include <vector>

using namespace fm;

namespace
{

struct Cursor
{
	const unsigned char* p;
	size_t n, at;
	bool ok;
	bool need(size_t k) { if (at + k > n) ok = false; return ok; }
	uint16_t u16() { if (!need(2)) return 0; uint16_t v = (uint16_t)((p[at] << 8) | p[at + 1]); at += 2; return v; }
	int32_t i32()
	{
		if (!need(4)) return 0;
		uint32_t v = ((uint32_t)p[at] << 24) | ((uint32_t)p[at + 1] << 16) | ((uint32_t)p[at + 2] << 8) | (uint32_t)p[at + 3];
		at += 4;
		return (int32_t)v;
	}
	void skip(size_t k) { if (need(k)) at += k; }
};

// header + point attribute table (BGEO.cpp:200-246, getAttributes :83-148); `bytes` must hold the start of the
// (uncompressed) file.  Returns FR_OK, or FR_ERR_STATE when more bytes are needed (info->data_offset = bytes wanted).
int parse_header(const unsigned char* bytes, size_t n, fr_bgeo_info* info)
{
	Cursor c{ bytes, n, 0, true };
	if (n < 5 || memcmp(bytes, "Bgeo", 4) != 0)
	{
		if (n >= 4 && bytes[0] == 0x7f && bytes[1] == 'N' && bytes[2] == 'S' && bytes[3] == 'J')
			set_error("bgeo: this is the new (JSON) bgeo format; only the classic format is supported, as in partio");
		else set_error("bgeo: magic number does not match 'Bgeo'");
		return FR_ERR_INVALID;
	}
	c.skip(5);                                    // magic + version char
	int32_t const version = c.i32();
	int32_t const n_points = c.i32();
	c.i32(); c.i32(); c.i32();                    // nPrims, nPointGroups, nPrimGroups
	int32_t const n_point_attr = c.i32();
	c.i32(); c.i32(); c.i32();                    // nVertexAttrib, nPrimAttrib, nAttrib
	if (!c.ok) { info->data_offset = 64; return FR_ERR_STATE; }
	if (version != 5) { set_error("bgeo: version must be 5"); return FR_ERR_INVALID; }
	if (n_points < 0 || n_point_attr < 0 || n_point_attr > 4096) { set_error("bgeo: corrupt header"); return FR_ERR_INVALID; }
	uint32_t words = 4;                           // x y z w
	for (int a = 0; a < n_point_attr; a++)
	{
		uint16_t const len = c.u16();
		c.skip(len);
		uint16_t const size = c.u16();
		int32_t const type = c.i32();
		if (!c.ok) break;
		if (type == 0 || type == 1 || type == 5) c.skip(4u * size);
		else if (type == 4)
		{
			int32_t const count = c.i32();
			if (count < 0) { set_error("bgeo: corrupt indexed-string attribute"); return FR_ERR_INVALID; }
			for (int32_t k = 0; k < count && c.ok; k++) c.skip(c.u16());
		}
		else { set_error("bgeo: unsupported point attribute type (partio aborts on it too)"); return FR_ERR_INVALID; }
		words += size;
	}
	if (!c.ok) { info->data_offset = n * 2 + 4096; return FR_ERR_STATE; }
	info->num_particles = (uint64_t)n_points;
	info->record_words = words;
	info->data_offset = c.at;
	return FR_OK;
}

int read_whole_file(const char* path, std::vector<unsigned char>& out)
{
	int const fd = open(path, O_RDONLY);
	if (fd < 0) { set_error(std::string("bgeo: cannot open ") + path + ": " + strerror(errno)); return FR_ERR_INVALID; }
	struct stat st;
	if (fstat(fd, &st) != 0) { close(fd); set_error("bgeo: fstat failed"); return FR_ERR_INVALID; }
	out.resize((size_t)st.st_size);
	size_t got = 0;
	while (got < out.size())
	{
		ssize_t const r = read(fd, out.data() + got, out.size() - got);
		if (r <= 0) break;
		got += (size_t)r;
	}
	close(fd);
	if (got != out.size()) { set_error("bgeo: short read"); return FR_ERR_INVALID; }
	return FR_OK;
}

// gzip member -> bytes (partio inflates .bgeo files that start with 1f 8b, ZIP.cpp:100-110, 260-300).  `want` > 0:
// stop after that many bytes (header probe).  The output is bounded: deflate cannot expand beyond ~1032 : 1, and
// FR_BGEO_MAX_INFLATED (default 16 GiB) caps what a crafted stream may make the host allocate.
int gunzip(const unsigned char* src, size_t n, std::vector<unsigned char>& out, size_t want = 0)
{
	static size_t const hard_cap = [] {
		const char* e = getenv("FR_BGEO_MAX_INFLATED");
		unsigned long long const v = e ? strtoull(e, nullptr, 10) : 0ull;
		return (size_t)(v ? v : (16ull << 30));
	}();
	z_stream zs;
	memset(&zs, 0, sizeof zs);
	if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) { set_error("bgeo: inflateInit2 failed"); return FR_ERR_INVALID; }
	size_t const first = want ? want : n * 4 + 65536;
	out.resize(first < hard_cap ? first : hard_cap);
	size_t consumed = 0, produced = 0;
	for (;;)
	{
		if (want && produced >= want) break;
		if (produced == out.size())
		{
			if (out.size() >= hard_cap)
			{
				inflateEnd(&zs);
				set_error("bgeo: gzip stream inflates beyond FR_BGEO_MAX_INFLATED");
				return FR_ERR_INVALID;
			}
			size_t const grown = out.size() * 2;
			out.resize(grown < hard_cap ? grown : hard_cap);
		}
		if (zs.avail_in == 0 && consumed < n)
		{
			size_t const chunk = n - consumed > 0x40000000u ? 0x40000000u : n - consumed;      // avail_in is 32 bits
			zs.next_in = const_cast<unsigned char*>(src) + consumed;
			zs.avail_in = (uInt)chunk;
			consumed += chunk;
		}
		zs.next_out = out.data() + produced;
		size_t const room = out.size() - produced;
		zs.avail_out = (uInt)(room > 0x40000000u ? 0x40000000u : room);
		uInt const before = zs.avail_out;
		int const rc = inflate(&zs, Z_NO_FLUSH);
		produced += before - zs.avail_out;
		if (rc == Z_STREAM_END) break;
		if (rc != Z_OK || (zs.avail_in == 0 && consumed == n && before == zs.avail_out))
		{
			inflateEnd(&zs);
			set_error("bgeo: corrupt gzip stream");
			return FR_ERR_INVALID;
		}
	}
	inflateEnd(&zs);
	out.resize(produced);
	return FR_OK;
}

// the file's uncompressed bytes + where the point block is
struct LoadedFile
{
	std::vector<unsigned char> raw, inflated;
	const unsigned char* bytes = nullptr;
	size_t size = 0;
	fr_bgeo_info info{};
};

// header_only: a gzip'd file is inflated only as far as its header and attribute table reach (fr_bgeo_probe); the
// length of the point block is then checked against the member's ISIZE trailer instead of the inflated bytes
int load_file(const char* path, LoadedFile& lf, bool header_only = false)
{
	if (!path) { set_error("bgeo: null path"); return FR_ERR_INVALID; }
	int rc = read_whole_file(path, lf.raw);
	if (rc) return rc;
	lf.bytes = lf.raw.data();
	lf.size = lf.raw.size();
	memset(&lf.info, 0, sizeof lf.info);
	bool const gz = lf.size >= 2 && lf.raw[0] == 0x1f && lf.raw[1] == 0x8b;
	if (gz && header_only)
	{
		fr_bgeo_info info;
		for (size_t want = 65536;; want *= 4)
		{
			if ((rc = gunzip(lf.raw.data(), lf.raw.size(), lf.inflated, want))) return rc;
			memset(&info, 0, sizeof info);
			rc = parse_header(lf.inflated.data(), lf.inflated.size(), &info);
			if (rc != FR_ERR_STATE || lf.inflated.size() < want) break;      // parsed, malformed, or the whole stream is in
		}
		if (rc == FR_ERR_STATE) { set_error("bgeo: file ends inside the header"); return FR_ERR_INVALID; }
		if (rc) return rc;
		info.compressed = 1;
		info.file_bytes = lf.raw.size();
		uint64_t const need = info.data_offset + info.num_particles * info.record_words * 4ull;
		uint32_t isize = 0;
		if (lf.raw.size() >= 18) memcpy(&isize, lf.raw.data() + lf.raw.size() - 4, 4);      // uncompressed size mod 2^32 (little-endian hosts)
		if (need < 0xffffffffull && lf.inflated.size() < need && (uint64_t)isize < need)
		{ set_error("bgeo: file ends inside the point block"); return FR_ERR_INVALID; }
		lf.info = info;
		lf.bytes = lf.inflated.data();
		lf.size = lf.inflated.size();
		return FR_OK;
	}
	if (gz)
	{
		if ((rc = gunzip(lf.raw.data(), lf.raw.size(), lf.inflated))) return rc;
		lf.bytes = lf.inflated.data();
		lf.size = lf.inflated.size();
		lf.info.compressed = 1;
	}
	rc = parse_header(lf.bytes, lf.size, &lf.info);
	if (rc == FR_ERR_STATE) { set_error("bgeo: file ends inside the header"); return FR_ERR_INVALID; }
	if (rc) return rc;
	lf.info.file_bytes = lf.raw.size();
	uint64_t const need = lf.info.data_offset + lf.info.num_particles * lf.info.record_words * 4ull;
	if (need > lf.size) { set_error("bgeo: file ends inside the point block"); return FR_ERR_INVALID; }
	return FR_OK;
}

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0u, 0x0123u); }

// point block (big-endian records of `words` 32-bit words, the first three are x y z) -> packed little-endian xyz.
// `raw` starts `shift` bytes after a 4-byte boundary (the header length is arbitrary): words are assembled from the
// two aligned words they straddle.
__global__ void __launch_bounds__(256) k_bgeo_unpack(const uint32_t* __restrict__ raw, uint32_t shift, uint32_t words, uint32_t n,
													 float* __restrict__ xyz)
{
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	size_t const w0 = (size_t)i * words;
	uint32_t v[3];
	if (shift == 0u)
	{
#pragma unroll
		for (int k = 0; k < 3; k++) v[k] = bswap32(__ldg(raw + w0 + k));
	}
	else
	{
		uint32_t a[4];
#pragma unroll
		for (int k = 0; k < 4; k++) a[k] = __ldg(raw + w0 + k);
		uint32_t const s = shift * 8u;
#pragma unroll
		for (int k = 0; k < 3; k++) v[k] = bswap32(__funnelshift_r(a[k], a[k + 1], s));    // bytes shift.. of the pair, little-endian memory
	}
	xyz[3ull * i] = __uint_as_float(v[0]);
	xyz[3ull * i + 1] = __uint_as_float(v[1]);
	xyz[3ull * i + 2] = __uint_as_float(v[2]);
}

int ensure_pinned(Context* c, size_t bytes)
{
	if (bytes <= c->cap_stage && c->h_stage) return FR_OK;
	if (c->h_stage) { cudaFreeHost(c->h_stage); c->h_stage = nullptr; c->cap_stage = 0; }
	size_t const want = bytes + bytes / 8 + 4096;
	FM_CUDA(cudaMallocHost((void**)&c->h_stage, want));
	c->cap_stage = want;
	return FR_OK;
}

}  // namespace

namespace fm
{

// the point block of `path` on the device as packed xyz in ctx->d_xyz; *n_out particles.  Work is left on the stream.
int stage_bgeo(Context* ctx, const char* path, size_t* n_out)
{
	// uncompressed files are read straight into pinned memory; gzip'd ones are inflated first
	int const fd = path ? open(path, O_RDONLY) : -1;
	if (fd < 0) { set_error(std::string("bgeo: cannot open ") + (path ? path : "(null)") + ": " + strerror(errno)); return FR_ERR_INVALID; }
	struct stat st;
	if (fstat(fd, &st) != 0 || st.st_size < 5) { close(fd); set_error("bgeo: empty or unreadable file"); return FR_ERR_INVALID; }
	size_t const fsize = (size_t)st.st_size;
	int rc = ensure_pinned(ctx, fsize + 16);
	if (rc) { close(fd); return rc; }
	// the previous upload from this staging buffer must have left it
	{ int const src = stream_sync(ctx); if (src) { close(fd); return src; } }
	size_t got = 0;
	while (got < fsize)
	{
		ssize_t const r = read(fd, ctx->h_stage + got, fsize - got);
		if (r <= 0) break;
		got += (size_t)r;
	}
	close(fd);
	if (got != fsize) { set_error("bgeo: short read"); return FR_ERR_INVALID; }
	const unsigned char* bytes = ctx->h_stage;
	size_t size = fsize;
	fr_bgeo_info info;
	memset(&info, 0, sizeof info);
	if (bytes[0] == 0x1f && bytes[1] == 0x8b)
	{
		std::vector<unsigned char> inflated;
		if ((rc = gunzip(bytes, size, inflated))) return rc;
		if ((rc = ensure_pinned(ctx, inflated.size() + 16))) return rc;
		memcpy(ctx->h_stage, inflated.data(), inflated.size());
		bytes = ctx->h_stage;
		size = inflated.size();
	}
	rc = parse_header(bytes, size, &info);
	if (rc == FR_ERR_STATE) { set_error("bgeo: file ends inside the header"); return FR_ERR_INVALID; }
	if (rc) return rc;
	if (info.num_particles == 0) { set_error("bgeo: file holds no particles"); return FR_ERR_INVALID; }
	if (info.num_particles > 0x7fffffffull) { set_error("bgeo: too many particles"); return FR_ERR_INVALID; }
	size_t const block = (size_t)info.num_particles * info.record_words * 4;
	if (info.data_offset + block > size) { set_error("bgeo: file ends inside the point block"); return FR_ERR_INVALID; }
	// copy from the 4-byte boundary at or before the block, so the device words stay aligned
	size_t const start = (size_t)info.data_offset & ~(size_t)3;
	uint32_t const shift = (uint32_t)(info.data_offset - start);
	size_t const copy_bytes = ((info.data_offset + block + 3) & ~(size_t)3) - start + 4;
	if ((rc = ensure_capacity(&ctx->d_raw, &ctx->cap_raw, copy_bytes / 4 + 1))) return rc;
	if ((rc = ensure_capacity(&ctx->d_xyz, &ctx->cap_xyz, (size_t)info.num_particles * 3))) return rc;
	FM_CUDA(cudaMemcpyAsync(ctx->d_raw, ctx->h_stage + start, copy_bytes <= ctx->cap_stage - start ? copy_bytes : ctx->cap_stage - start,
							cudaMemcpyHostToDevice, ctx->stream));
	uint32_t const n32 = (uint32_t)info.num_particles;
	k_bgeo_unpack<<<(n32 + 255u) / 256u, 256, 0, ctx->stream>>>(ctx->d_raw, shift, info.record_words, n32, ctx->d_xyz);
	ctx->kernel_launches += 1;
	FM_CUDA(cudaGetLastError());
	*n_out = (size_t)info.num_particles;
	return FR_OK;
}

}  // namespace fm

extern "C" {

int fr_bgeo_probe(const char* path, fr_bgeo_info* out)
{
	if (!out) { set_error("fr_bgeo_probe: null out"); return FR_ERR_INVALID; }
	LoadedFile lf;
	int const rc = load_file(path, lf, true);
	if (rc) return rc;
	*out = lf.info;
	return FR_OK;
}

int fr_bgeo_read(const char* path, float* xyz, uint64_t capacity, uint64_t* n)
{
	LoadedFile lf;
	int const rc = load_file(path, lf);
	if (rc) return rc;
	if (n) *n = lf.info.num_particles;
	if (!xyz) return FR_OK;
	if (capacity < lf.info.num_particles) { set_error("fr_bgeo_read: output array too small"); return FR_ERR_INVALID; }
	const unsigned char* p = lf.bytes + lf.info.data_offset;
	size_t const stride = (size_t)lf.info.record_words * 4;
	for (uint64_t i = 0; i < lf.info.num_particles; i++, p += stride)
		for (int k = 0; k < 3; k++)
		{
			uint32_t const v = ((uint32_t)p[4 * k] << 24) | ((uint32_t)p[4 * k + 1] << 16) | ((uint32_t)p[4 * k + 2] << 8) | (uint32_t)p[4 * k + 3];
			memcpy(xyz + 3 * i + k, &v, 4);
		}
	return FR_OK;
}

// positions only, the bytes partio's writeBGEO produces for a particle set whose only attribute is "position"
// (BGEO.cpp:304-460): header, records (x, y, z, w = 1), no primitives, no detail attributes, 0x00 0xff
int fr_bgeo_write(const char* path, const float* xyz, uint64_t n, int compressed)
{
	if (!path || (!xyz && n) || n > 0x7fffffffull) { set_error("fr_bgeo_write: bad arguments"); return FR_ERR_INVALID; }
	std::vector<unsigned char> out;
	out.reserve(64 + (size_t)n * 16);
	auto put32 = [&](uint32_t v) { out.push_back((unsigned char)(v >> 24)); out.push_back((unsigned char)(v >> 16)); out.push_back((unsigned char)(v >> 8)); out.push_back((unsigned char)v); };
	out.insert(out.end(), { 'B', 'g', 'e', 'o', 'V' });
	put32(5); put32((uint32_t)n);
	for (int k = 0; k < 7; k++) put32(0);         // nPrims, nPointGroups, nPrimGroups, nPointAttrib, nVertexAttrib, nPrimAttrib, nAttrib
	for (uint64_t i = 0; i < n; i++)
	{
		for (int k = 0; k < 3; k++) { uint32_t v; memcpy(&v, xyz + 3 * i + k, 4); put32(v); }
		put32(0x3f800000u);
	}
	out.push_back(0x00); out.push_back(0xff);
	if (compressed)
	{
		gzFile g = gzopen(path, "wb");
		if (!g) { set_error(std::string("fr_bgeo_write: cannot create ") + path); return FR_ERR_INVALID; }
		size_t done = 0;
		while (done < out.size())
		{
			unsigned const chunk = (unsigned)((out.size() - done) > (1u << 30) ? (1u << 30) : (out.size() - done));
			if (gzwrite(g, out.data() + done, chunk) != (int)chunk) { gzclose(g); set_error("fr_bgeo_write: gzwrite failed"); return FR_ERR_INVALID; }
			done += chunk;
		}
		gzclose(g);
		return FR_OK;
	}
	int const fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
	if (fd < 0) { set_error(std::string("fr_bgeo_write: cannot create ") + path + ": " + strerror(errno)); return FR_ERR_INVALID; }
	size_t done = 0;
	while (done < out.size())
	{
		ssize_t const w = write(fd, out.data() + done, out.size() - done);
		if (w <= 0) { close(fd); set_error("fr_bgeo_write: write failed"); return FR_ERR_INVALID; }
		done += (size_t)w;
	}
	close(fd);
	return FR_OK;
}

// Dataset::Dataset's enumeration (Dataset.cpp:186-203): <prefix><i><suffix>, i = 1, 2, ... until a file is missing or
// `count` (< 0: unbounded) is reached
int fr_dataset_count(const char* prefix, const char* suffix, int count)
{
	if (!prefix || !suffix) { set_error("fr_dataset_count: null argument"); return FR_ERR_INVALID; }
	int i = 1;
	while (count < 0 || i <= count)
	{
		std::string const path = std::string(prefix) + std::to_string(i) + suffix;
		if (access(path.c_str(), R_OK) != 0) break;
		++i;
	}
	return i - 1;
}

int fr_upload_frame_bgeo(fr_context* ctx, int frame, const char* path, float h, float h_ext_mult)
{
	if (!ctx) { set_error("null context"); return FR_ERR_INVALID; }
	FM_CUDA(cudaSetDevice(ctx->device));
	int rc = fr_wait(ctx);             // a pending render's events are read before they are recorded again
	if (rc) return rc;
	FM_TIME(ctx, ctx->ev[0], ctx->stream);
	size_t n = 0;
	if ((rc = stage_bgeo(ctx, path, &n))) return rc;
	FM_TIME(ctx, ctx->ev[1], ctx->stream);
	rc = fr_build_frame_device(ctx, frame, ctx->d_xyz, n, h, h_ext_mult);
	if (rc == FR_OK && ctx->build_timed) ctx->build_timed = 2;      // upload_ms = file block copy + decode
	return rc;
}

}  // extern "C"
