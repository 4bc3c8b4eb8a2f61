// Host side of the reference's own call shape: Stitcher.stitch(images_dic) hands over numpy frames,
// i.e. PAGEABLE host memory (StitcherClass.py:114-136, :239).  cudaMemcpy from pageable memory is staged
// by the driver on the calling thread, one copy after the other (1.6 ms for the 24.9 MB config 2's six
// cameras have to send).  mcs_upload_pageable_u8 does that staging itself: a small persistent pool of
// host threads copies the windows into the caller's pinned staging frames in pieces, and the calling
// thread issues the DMA of every piece as soon as it has landed.  No Python, no GIL in the loop.
#include "mcs_common.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <exception>
#include <memory>
#include <mutex>
#include <string.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace {

struct Piece {
    char* stage;         // first byte of the piece in the pinned staging frame
    const char* src;     // ... in the caller's pageable frame
    char* dst;           // ... in the device frame
    int64_t pitch;       // row pitch of staging and device frame
    int64_t src_pitch;
    int64_t width;
    int rows;
};

// One call = one Job.  Workers take a reference to the job that is current when they wake up, claim its pieces
// with the job's own cursor and flag them done; a worker that wakes up late finds a finished (or no) job and
// goes back to sleep, it can never claim a piece of a later call.
struct Job {
    std::vector<Piece> pieces;
    std::unique_ptr<std::atomic<char>[]> done;
    std::atomic<int> cursor{0};
    std::atomic<int> joined{0};
    int threads = 1;
    std::mutex mu;                 // the issuing thread sleeps here while the piece it waits for is being copied
    std::condition_variable landed;
};

class StagePool {
public:
    static StagePool& get() {
        static StagePool* p = new StagePool();   // never destroyed: the workers are detached from static destruction
        return *p;
    }

    cudaError_t run(std::vector<Piece>&& pieces, int threads, cudaStream_t stream) {
        std::lock_guard<std::mutex> call(call_mu_);    // one job at a time
        auto job = std::make_shared<Job>();
        job->pieces = std::move(pieces);
        const int n = (int)job->pieces.size();
        job->done.reset(new std::atomic<char>[n]);
        for (int i = 0; i < n; ++i) job->done[i].store(0, std::memory_order_relaxed);
        job->threads = std::max(1, std::min(threads, MAX_THREADS));
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (pid_ != getpid()) {      // a forked child inherits the count, not the threads
                pid_ = getpid();
                n_workers_ = 0;
            }
            for (; n_workers_ < job->threads; ++n_workers_) std::thread([this] { worker(); }).detach();
            job_ = job;
            ++generation_;
        }
        cv_.notify_all();
        cudaError_t err = cudaSuccess;
        for (int i = 0; i < n; ++i) {
            if (!job->done[i].load(std::memory_order_acquire)) {
                std::unique_lock<std::mutex> lk(job->mu);
                job->landed.wait(lk, [&] { return job->done[i].load(std::memory_order_acquire) != 0; });
            }
            if (err == cudaSuccess) {
                const Piece& p = job->pieces[i];
                err = cudaMemcpy2DAsync(p.dst, (size_t)p.pitch, p.stage, (size_t)p.pitch, (size_t)p.width, (size_t)p.rows,
                                        cudaMemcpyHostToDevice, stream);
            }
        }
        std::lock_guard<std::mutex> lk(mu_);
        job_.reset();
        return err;
    }

private:
    static constexpr int MAX_THREADS = 16;

    void worker() {
        unsigned long long seen = 0;
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                job = job_;
            }
            if (!job || job->joined.fetch_add(1, std::memory_order_relaxed) >= job->threads) continue;
            const int n = (int)job->pieces.size();
            for (;;) {
                const int i = job->cursor.fetch_add(1, std::memory_order_relaxed);
                if (i >= n) break;
                const Piece& p = job->pieces[i];
                if (p.pitch == p.width && p.src_pitch == p.width) {
                    memcpy(p.stage, p.src, (size_t)p.width * (size_t)p.rows);
                } else {
                    for (int r = 0; r < p.rows; ++r)
                        memcpy(p.stage + (size_t)r * p.pitch, p.src + (size_t)r * p.src_pitch, (size_t)p.width);
                }
                {
                    std::lock_guard<std::mutex> lk(job->mu);
                    job->done[i].store(1, std::memory_order_release);
                }
                job->landed.notify_one();
            }
        }
    }

    std::mutex call_mu_, mu_;
    std::condition_variable cv_;
    int n_workers_ = 0;
    pid_t pid_ = 0;
    std::shared_ptr<Job> job_;
    unsigned long long generation_ = 0;
};

}  // namespace

extern "C" int mcs_upload_pageable_u8(int n_windows, void* const* dst, const void* const* src, void* const* staging,
                                      const int64_t* pitch_bytes, const int64_t* src_pitch_bytes, const int64_t* xywh,
                                      int64_t piece_bytes, int threads, void* cuda_stream) {
    MCS_CHECK_ARG(n_windows >= 0, "mcs_upload_pageable_u8: negative window count");
    if (n_windows == 0) return MCS_OK;
    MCS_CHECK_ARG(dst && src && staging && pitch_bytes && src_pitch_bytes && xywh, "mcs_upload_pageable_u8: NULL table");
    MCS_CHECK_ARG(threads >= 1, "mcs_upload_pageable_u8: needs at least one staging thread");
    if (piece_bytes <= 0) piece_bytes = 1 << 20;
    try {   // std::vector / std::thread may throw; nothing crosses the C boundary
    std::vector<Piece> pieces;
    for (int i = 0; i < n_windows; ++i) {
        const int64_t x0 = xywh[4 * i], y0 = xywh[4 * i + 1], w = xywh[4 * i + 2], h = xywh[4 * i + 3];
        MCS_CHECK_ARG(dst[i] && src[i] && staging[i], "mcs_upload_pageable_u8: NULL buffer in window %d", i);
        MCS_CHECK_ARG(x0 >= 0 && y0 >= 0 && w >= 0 && h >= 0, "mcs_upload_pageable_u8: negative extent in window %d", i);
        MCS_CHECK_ARG(x0 + w <= pitch_bytes[i] && x0 + w <= src_pitch_bytes[i],
                      "mcs_upload_pageable_u8: window %d wider than a row", i);
        if (w == 0 || h == 0) continue;
        const int64_t step = std::max<int64_t>(8, piece_bytes / w);
        for (int64_t y = y0; y < y0 + h; y += step) {
            Piece p;
            const size_t off = (size_t)y * (size_t)pitch_bytes[i] + (size_t)x0;
            p.stage = static_cast<char*>(staging[i]) + off;
            p.dst = static_cast<char*>(dst[i]) + off;
            p.src = static_cast<const char*>(src[i]) + (size_t)y * (size_t)src_pitch_bytes[i] + (size_t)x0;
            p.pitch = pitch_bytes[i];
            p.src_pitch = src_pitch_bytes[i];
            p.width = w;
            p.rows = (int)std::min<int64_t>(step, y0 + h - y);
            pieces.push_back(p);
        }
    }
    if (pieces.empty()) return MCS_OK;
    const cudaError_t err = StagePool::get().run(std::move(pieces), threads, (cudaStream_t)cuda_stream);
    if (err != cudaSuccess) {
        mcs_set_error("mcs_upload_pageable_u8: cudaMemcpy2DAsync failed: %s", cudaGetErrorString(err));
        return MCS_ERR_CUDA;
    }
    return MCS_OK;
    } catch (const std::exception& e) {
        mcs_set_error("mcs_upload_pageable_u8: %s", e.what());
        return MCS_ERR_NOMEM;
    }
}
