// fm_record.cu -- recording: the finished colour image as the .bmp the reference's screenshot path writes.
//
// Replaces Renderer::_Screenshot (src/engine/renderer/Renderer.cpp:326-415): swapchain image -> host-visible linear
// image -> a host loop that swaps the R and B bytes of every pixel -> stbi_write_bmp(filename, W, H, 4, data)
// (vendor/stb_image/stb_image_write.h:492-510: 14-byte file header + 108-byte BITMAPV4HEADER, BI_BITFIELDS, 32 bpp,
// rows bottom-up, pixels B G R A), called once per recorded frame (AdvancedRenderer.cpp:283-297,
// tools/screenshots_to_video.ps1).  Here the swizzle and the vertical flip are a kernel on the image that is already
// on the device; the host only prepends the header and writes the file, on the lane's worker when part of a sequence.
#include "fm_internal.h"

#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <unistd.h>

#include <string>

using namespace fm;

namespace
{

constexpr size_t kBmpHeader = 14 + 108;

// R G B A rows top-down -> B G R A rows bottom-up
__global__ void __launch_bounds__(256) k_bmp_pack(const uchar4* __restrict__ rgba, int W, int H, uchar4* __restrict__ out)
{
	int const x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
	if (x >= W) return;
	uchar4 const p = rgba[(size_t)y * W + x];
	out[(size_t)(H - 1 - y) * W + x] = make_uchar4(p.z, p.y, p.x, p.w);
}

void put16(unsigned char*& p, uint32_t v) { *p++ = (unsigned char)v; *p++ = (unsigned char)(v >> 8); }
void put32(unsigned char*& p, uint32_t v) { put16(p, v & 0xffffu); put16(p, v >> 16); }

// stbi_write_bmp_core, comp == 4 (stb_image_write.h:501-509)
void bmp_header(unsigned char* p, int W, int H)
{
	*p++ = 'B'; *p++ = 'M';
	put32(p, (uint32_t)(kBmpHeader + (size_t)W * H * 4)); put16(p, 0); put16(p, 0); put32(p, (uint32_t)kBmpHeader);
	put32(p, 108); put32(p, (uint32_t)W); put32(p, (uint32_t)H); put16(p, 1); put16(p, 32);
	put32(p, 3); put32(p, 0); put32(p, 0); put32(p, 0); put32(p, 0); put32(p, 0);
	put32(p, 0xff0000u); put32(p, 0xff00u); put32(p, 0xffu); put32(p, 0xff000000u);
	put32(p, 0);
	for (int k = 0; k < 12; k++) put32(p, 0);
}

}  // namespace

extern "C" {

int fr_encode_bmp(fr_context* ctx, uint8_t* out, size_t capacity, size_t* bytes)
{
	if (!ctx) { set_error("null context"); return FR_ERR_INVALID; }
	FM_CUDA(cudaSetDevice(ctx->device));
	int const W = ctx->width, H = ctx->height;
	size_t const need = kBmpHeader + (size_t)W * H * 4;
	if (bytes) *bytes = need;
	if (!out) return FR_OK;
	if (capacity < need) { set_error("fr_encode_bmp: output buffer too small"); return FR_ERR_INVALID; }
	int rc = fr_wait(ctx);
	if (rc) return rc;
	if ((rc = ensure_capacity(&ctx->d_bmp, &ctx->cap_bmp, (size_t)W * H))) return rc;
	dim3 const grid((W + 255) / 256, H);
	k_bmp_pack<<<grid, 256, 0, ctx->stream>>>(ctx->d_rgba_target, W, H, ctx->d_bmp);
	ctx->kernel_launches += 1;
	FM_CUDA(cudaGetLastError());
	FM_CUDA(cudaMemcpyAsync(out + kBmpHeader, ctx->d_bmp, (size_t)W * H * 4, cudaMemcpyDeviceToHost, ctx->stream));
	bmp_header(out, W, H);
	return stream_sync(ctx);
}

int fr_write_bmp(fr_context* ctx, const char* path)
{
	if (!ctx || !path) { set_error("fr_write_bmp: null argument"); return FR_ERR_INVALID; }
	FM_CUDA(cudaSetDevice(ctx->device));
	size_t const need = kBmpHeader + (size_t)ctx->width * ctx->height * 4;
	if (need > ctx->cap_bmp_host || !ctx->h_bmp)
	{
		if (ctx->h_bmp) { cudaFreeHost(ctx->h_bmp); ctx->h_bmp = nullptr; ctx->cap_bmp_host = 0; }
		FM_CUDA(cudaMallocHost((void**)&ctx->h_bmp, need));
		ctx->cap_bmp_host = need;
	}
	int const rc = fr_encode_bmp(ctx, ctx->h_bmp, ctx->cap_bmp_host, nullptr);
	if (rc) return rc;
	int const fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
	if (fd < 0) { set_error(std::string("fr_write_bmp: cannot create ") + path + ": " + strerror(errno)); return FR_ERR_INVALID; }
	size_t done = 0;
	while (done < need)
	{
		ssize_t const w = write(fd, ctx->h_bmp + done, need - done);
		if (w <= 0) { close(fd); set_error("fr_write_bmp: write failed"); return FR_ERR_INVALID; }
		done += (size_t)w;
	}
	close(fd);
	return FR_OK;
}

}  // extern "C"
