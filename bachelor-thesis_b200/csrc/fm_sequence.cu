// fm_sequence.cu -- frame pipeline over several contexts of one GPU (fr_seq_* of include/fluidmarch.h).
//
// The reference renders an animation one frame at a time: AdvancedRenderer::Render starts a march, polls IsDone once
// per UI frame and only then advances `Frame` (src/app/AdvancedRenderer/AdvancedRenderer.cpp:257-298); its only
// parallelism is the pixel ThreadPool (src/app/ThreadPool.cpp:38-55).  On the GPU a single frame leaves gaps no kernel
// can fill by itself: the host round trip that sizes the grid tables, one-CTA scan steps, the tail of the persistent
// march where the last tiles and the few long rays run on a mostly idle machine, and -- end to end -- the PCIe copies
// of the particles in and the image out.  A sequence keeps `lanes` frames in flight instead: one fr_context (own
// stream, own images and scratch) and one host worker thread per lane; frame k goes to lane k % lanes.  Kernels of
// different lanes overlap wherever SMs are free and the copy engines run beside them, so the gaps of one frame are
// filled with the next frame's work.  Every frame is still rendered by exactly the single-context code path, hence
// bit-identical to fr_render_async on one context (tests/test_gpu_sequence.py).
#include "fm_internal.h"

#include <stdlib.h>

#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <utility>
#include <vector>

using namespace fm;

namespace
{

struct Lane
{
	fr_context* ctx = nullptr;
	std::thread worker;
	std::mutex m;
	std::condition_variable cv;
	bool has_job = false, busy = false, quit = false;
	fr_seq_job job{};
	std::string path, bmp;         // copies of job.bgeo_path, job.bmp_path
	int64_t ticket = -1;           // ticket of the job in `job` / being worked on
	int64_t done_ticket = -1;      // last ticket finished on this lane
	int done_status = FR_OK;
	std::string done_error;
	// status of the frames that FAILED on this lane and have not been waited for yet: a caller that submits ahead may
	// ask for ticket t after ticket t + lanes has already finished here (fr_seq_wait; cleared by wait and drain)
	std::vector<std::pair<int64_t, std::pair<int, std::string>>> failed;
	cudaEvent_t ev_end = nullptr;
	// A frame as two CUDA graphs, each re-captured every frame and patched into its instantiated graph: (a) upload,
	// frame build, depth pre-pass; (b) march, copies out.  Between them the worker polls the grid parameters in mapped
	// memory.  Two launches instead of ~30 stream operations, which is what a host -> host sequence is short of while
	// bulk copies keep PCIe busy (every stream operation gets slower then, tools/e2e_probe2.py)
	cudaGraphExec_t exec = nullptr, exec_a = nullptr;
	bool graph_ok = true;
};

}  // namespace

struct fr_sequence
{
	int device = 0;
	std::vector<Lane*> lanes;
	int64_t next_ticket = 0;
	int first_error = FR_OK;
	std::string first_error_text;
	std::mutex err_m;
	cudaEvent_t ev_begin = nullptr;
	uint64_t frames_done = 0;
	bool graphs = true;
};

namespace
{

// one part of the frame through a CUDA graph (see Lane::exec).  Any failure of the capture machinery switches the lane
// back to plain launches; the part itself is then enqueued directly.
template <typename Fn>
int enqueue_as_graph(Lane* ln, cudaGraphExec_t* exec, Fn&& enqueue)
{
	fr_context* const c = ln->ctx;
	if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess)
	{
		cudaGetLastError();
		ln->graph_ok = false;
		return enqueue();
	}
	int const rc = enqueue();
	cudaGraph_t g = nullptr;
	cudaError_t e = cudaStreamEndCapture(c->stream, &g);
	if (rc != FR_OK) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return rc; }      // a real error of the frame
	if (e == cudaSuccess && *exec)
	{
		cudaGraphExecUpdateResultInfo info;
		if (cudaGraphExecUpdate(*exec, g, &info) != cudaSuccess)       // topology changed (kernel variant, an extra memset)
		{
			cudaGetLastError();
			cudaGraphExecDestroy(*exec);
			*exec = nullptr;
		}
	}
	if (e == cudaSuccess && !*exec) e = cudaGraphInstantiate(exec, g, 0);
	if (e == cudaSuccess) e = cudaGraphLaunch(*exec, c->stream);
	if (g) cudaGraphDestroy(g);
	if (e != cudaSuccess)
	{
		// (nothing of this part has reached the stream: the capture only recorded it)
		cudaGetLastError();
		ln->graph_ok = false;
		if (*exec) { cudaGraphExecDestroy(*exec); *exec = nullptr; }
		c->render_pending = false;
		return enqueue();
	}
	return FR_OK;
}

void lane_main(fr_sequence* seq, Lane* ln)
{
	cudaSetDevice(seq->device);
	for (;;)
	{
		fr_seq_job job;
		std::string path, bmp;
		int64_t ticket;
		{
			std::unique_lock<std::mutex> lk(ln->m);
			ln->cv.wait(lk, [&] { return ln->has_job || ln->quit; });
			if (!ln->has_job && ln->quit) return;
			job = ln->job;
			path = ln->path;
			bmp = ln->bmp;
			ticket = ln->ticket;
			ln->has_job = false;
			ln->busy = true;
		}
		bool launches_only = false;
		int rc = lane_frame_begin(ln->ctx, job, job.bgeo_path ? path.c_str() : nullptr, &launches_only);
		if (rc == FR_OK) rc = lane_frame_upload(ln->ctx, job);
		if (rc == FR_OK)
			rc = seq->graphs && ln->graph_ok && launches_only
				? enqueue_as_graph(ln, &ln->exec_a, [&] { return lane_frame_stage_a(ln->ctx, job); })
				: lane_frame_stage_a(ln->ctx, job);
		if (rc == FR_OK) rc = lane_frame_resolve(ln->ctx, job);
		if (rc == FR_OK)
			rc = seq->graphs && ln->graph_ok ? enqueue_as_graph(ln, &ln->exec, [&] { return lane_frame_enqueue(ln->ctx, job); })
											 : lane_frame_enqueue(ln->ctx, job);
		if (rc == FR_OK) rc = lane_frame_wait(ln->ctx);       // the whole stream: build, render and copies
		if (rc == FR_OK && job.bmp_path) rc = fr_write_bmp(ln->ctx, bmp.c_str());      // recording (Renderer.cpp:400-409)
		std::string err;
		if (rc != FR_OK) err = fr_last_error();      // thread-local text of this worker
		{
			std::lock_guard<std::mutex> lk(ln->m);
			ln->busy = false;
			ln->done_ticket = ticket;
			ln->done_status = rc;
			ln->done_error = err;
			if (rc != FR_OK)
			{
				if (ln->failed.size() >= 64) ln->failed.erase(ln->failed.begin());
				ln->failed.push_back({ ticket, { rc, err } });
			}
		}
		if (rc != FR_OK)
		{
			std::lock_guard<std::mutex> lk(seq->err_m);
			if (seq->first_error == FR_OK) { seq->first_error = rc; seq->first_error_text = err; }
		}
		ln->cv.notify_all();
	}
}

int wait_lane_idle(Lane* ln)
{
	std::unique_lock<std::mutex> lk(ln->m);
	ln->cv.wait(lk, [&] { return !ln->has_job && !ln->busy; });
	return ln->done_status;
}

}  // namespace

extern "C" {

int fr_seq_create(int device, int width, int height, int lanes, fr_sequence** out)
{
	if (!out || lanes < 1 || lanes > 16) { set_error("fr_seq_create: lanes must be in [1, 16]"); return FR_ERR_INVALID; }
	*out = nullptr;
	fr_sequence* seq = new (std::nothrow) fr_sequence();
	if (!seq) { set_error("out of host memory"); return FR_ERR_INVALID; }
	seq->device = device;
	if (const char* e = getenv("FR_SEQ_GRAPH")) seq->graphs = e[0] != '0';
	for (int k = 0; k < lanes; k++)
	{
		Lane* ln = new (std::nothrow) Lane();
		if (!ln) { fr_seq_destroy(seq); set_error("out of host memory"); return FR_ERR_INVALID; }
		seq->lanes.push_back(ln);
		int const rc = fr_create(device, width, height, &ln->ctx);
		if (rc != FR_OK) { fr_seq_destroy(seq); return rc; }
		ln->ctx->stage_timing = false;       // no per-stage events in a sequence (see Context::stage_timing)
		{ const char* e = getenv("FLUIDMARCH_OVERLAP_LANES"); ln->ctx->overlap_depth = e && e[0] == '1'; }   // (the other lanes' frames fill the GPU)
		if (cudaEventCreate(&ln->ev_end) != cudaSuccess) { fr_seq_destroy(seq); return cuda_fail(cudaGetLastError(), "cudaEventCreate", __FILE__, __LINE__); }
	}
	if (cudaEventCreate(&seq->ev_begin) != cudaSuccess) { fr_seq_destroy(seq); return cuda_fail(cudaGetLastError(), "cudaEventCreate", __FILE__, __LINE__); }
	for (Lane* ln : seq->lanes) ln->worker = std::thread(lane_main, seq, ln);
	*out = seq;
	return FR_OK;
}

void fr_seq_destroy(fr_sequence* seq)
{
	if (!seq) return;
	for (Lane* ln : seq->lanes)
	{
		if (ln->worker.joinable())
		{
			{ std::lock_guard<std::mutex> lk(ln->m); ln->quit = true; }
			ln->cv.notify_all();
			ln->worker.join();
		}
		cudaSetDevice(seq->device);
		if (ln->ev_end) cudaEventDestroy(ln->ev_end);
		if (ln->exec) cudaGraphExecDestroy(ln->exec);
		if (ln->exec_a) cudaGraphExecDestroy(ln->exec_a);
		if (ln->ctx) fr_destroy(ln->ctx);
		delete ln;
	}
	if (seq->ev_begin) cudaEventDestroy(seq->ev_begin);
	delete seq;
}

int fr_seq_set_yielding(fr_sequence* seq, int on)
{
	int const rc = fr_seq_drain(seq);
	if (rc) return rc;
	for (Lane* ln : seq->lanes) ln->ctx->blocking_sync = on != 0;
	return FR_OK;
}

int fr_seq_lanes(fr_sequence* seq) { return seq ? (int)seq->lanes.size() : FR_ERR_INVALID; }

int fr_seq_context(fr_sequence* seq, int lane, fr_context** out)
{
	if (!seq || !out || lane < 0 || (size_t)lane >= seq->lanes.size()) { set_error("fr_seq_context: bad lane"); return FR_ERR_INVALID; }
	*out = seq->lanes[lane]->ctx;
	return FR_OK;
}

int fr_seq_drain(fr_sequence* seq)
{
	if (!seq) { set_error("null sequence"); return FR_ERR_INVALID; }
	for (Lane* ln : seq->lanes)
	{
		wait_lane_idle(ln);
		std::lock_guard<std::mutex> lk(ln->m);
		ln->failed.clear();
	}
	std::lock_guard<std::mutex> lk(seq->err_m);
	int const rc = seq->first_error;
	if (rc != FR_OK) set_error(seq->first_error_text);
	seq->first_error = FR_OK;
	seq->first_error_text.clear();
	return rc;
}

int fr_seq_set_camera(fr_sequence* seq, const fr_camera* cam)
{
	int rc = fr_seq_drain(seq);
	if (rc) return rc;
	for (Lane* ln : seq->lanes)
		if ((rc = fr_set_camera(ln->ctx, cam))) return rc;
	return FR_OK;
}

int fr_seq_set_settings(fr_sequence* seq, const fr_settings* s)
{
	int rc = fr_seq_drain(seq);
	if (rc) return rc;
	if (!s) { set_error("fr_seq_set_settings: null settings"); return FR_ERR_INVALID; }
	fr_settings t = *s;
	t.frame = 0;                     // every lane keeps its current frame in slot 0
	for (Lane* ln : seq->lanes)
		if ((rc = fr_set_settings(ln->ctx, &t))) return rc;
	return FR_OK;
}

int64_t fr_seq_submit(fr_sequence* seq, const fr_seq_job* job)
{
	if (!seq || !job || (!job->bgeo_path && (!job->xyz || job->n == 0))) { set_error("fr_seq_submit: bad job"); return FR_ERR_INVALID; }
	int64_t const ticket = seq->next_ticket++;
	Lane* ln = seq->lanes[(size_t)(ticket % (int64_t)seq->lanes.size())];
	{
		std::unique_lock<std::mutex> lk(ln->m);
		ln->cv.wait(lk, [&] { return !ln->has_job && !ln->busy; });    // the lane's previous frame (ticket - lanes) is out
		ln->job = *job;
		ln->path = job->bgeo_path ? job->bgeo_path : "";
		ln->bmp = job->bmp_path ? job->bmp_path : "";
		ln->ticket = ticket;
		ln->has_job = true;
	}
	ln->cv.notify_all();
	return ticket;
}

int fr_seq_wait(fr_sequence* seq, int64_t ticket)
{
	if (!seq || ticket < 0 || ticket >= seq->next_ticket) { set_error("fr_seq_wait: unknown ticket"); return FR_ERR_INVALID; }
	Lane* ln = seq->lanes[(size_t)(ticket % (int64_t)seq->lanes.size())];
	std::unique_lock<std::mutex> lk(ln->m);
	ln->cv.wait(lk, [&] { return ln->done_ticket >= ticket; });
	for (size_t k = 0; k < ln->failed.size(); k++)
		if (ln->failed[k].first == ticket)
		{
			int const rc = ln->failed[k].second.first;
			set_error(ln->failed[k].second.second);
			ln->failed.erase(ln->failed.begin() + (long)k);
			return rc;
		}
	return FR_OK;
}

// device time of everything submitted between the two calls: begin is recorded on lane 0's stream with every lane
// idle, end on every lane's stream as soon as the last frame has drained (the streams are idle by then, so the end
// stamps are late by the drain's wake-up latency, a few 10 us per measurement: the time errs on the long side; a
// per-frame end event would be exact but is one more stream operation per frame, see Context::stage_timing)
int fr_seq_timer_begin(fr_sequence* seq)
{
	int rc = fr_seq_drain(seq);
	if (rc) return rc;
	FM_CUDA(cudaSetDevice(seq->device));
	FM_CUDA(cudaDeviceSynchronize());
	FM_CUDA(cudaEventRecord(seq->ev_begin, seq->lanes[0]->ctx->stream));
	return FR_OK;
}

int fr_seq_timer_end(fr_sequence* seq, float* ms)
{
	int rc = fr_seq_drain(seq);
	if (rc) return rc;
	if (!ms) { set_error("fr_seq_timer_end: null out"); return FR_ERR_INVALID; }
	FM_CUDA(cudaSetDevice(seq->device));
	for (Lane* ln : seq->lanes) FM_CUDA(cudaEventRecord(ln->ev_end, ln->ctx->stream));
	float best = 0.0f;
	for (Lane* ln : seq->lanes)
	{
		FM_CUDA(cudaEventSynchronize(ln->ev_end));
		float t = 0.0f;
		FM_CUDA(cudaEventElapsedTime(&t, seq->ev_begin, ln->ev_end));
		if (t > best) best = t;
	}
	*ms = best;
	return FR_OK;
}

}  // extern "C"
