// frame_writer.hpp -- writes the frames of the timestep loop off the critical path.
//
// The reference's CUDA driver stops the device after every step, copies the float image back with a blocking cudaMemcpy
// and encodes the PNG on the same thread (main.cpp:228-232, :243-245; the cudaMemcpyAsync over NUM_STREAMS streams it
// gestures at is commented out).  Here a frame goes through a ring of slots:
//   compute stream : the bytes of the frame are formed in the slot's device buffer (utils.hpp:129-131): by the kernel that
//                    advects the image (pfs_ctx_advect_color_step_rgba8) or by pfs_image_to_rgba8
//   copy stream    : waits for that pack, copies the BYTES to the slot's pinned host buffer
//   worker thread  : waits for the copy, encodes <dir>/<i>.png with libpng, frees the slot
// so the next timestep starts while frame i is still being copied and encoded, and several frames encode in parallel.
// The files are the same bytes the serial path would write.
#pragma once

#include <condition_variable>
#include <cstdio>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime_api.h>

#include "../../include/pfs_b200.h"
#include "png_io.hpp"

class FrameWriter {
public:
    FrameWriter(const pngio::png_image &header, int width, int height, int channels, int n_slots, int n_workers)
        : header_(header), w_(width), h_(height), c_(channels), bytes_((size_t)width * height * channels)
    {
        header_.opaque = nullptr;
        ok_ = cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking) == cudaSuccess;
        slots_.resize(n_slots);
        for (auto &s : slots_) {
            ok_ = ok_ && cudaMalloc((void **)&s.dev, bytes_) == cudaSuccess &&
                  cudaMallocHost((void **)&s.host, bytes_) == cudaSuccess &&
                  cudaEventCreateWithFlags(&s.packed, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming) == cudaSuccess;
            free_.push_back(&s);
        }
        for (int i = 0; i < n_workers; i++) workers_.emplace_back([this] { work(); });
    }

    ~FrameWriter()
    {
        finish();
        for (auto &s : slots_) {
            if (s.dev) cudaFree(s.dev);
            if (s.host) cudaFreeHost(s.host);
            if (s.packed) cudaEventDestroy(s.packed);
            if (s.copied) cudaEventDestroy(s.copied);
        }
        if (copy_stream_) cudaStreamDestroy(copy_stream_);
    }

    bool ok() const { return ok_; }

    // Enqueue frame `path` from the float image on the device (legacy default stream, where the steps run).
    // Blocks only while every slot is still in flight.  false: a CUDA call failed (message in error()).
    bool submit(const float *d_image, const std::string &path)
    {
        unsigned char *dst = begin(path);
        if (pfs_image_to_rgba8(d_image, dst, w_, h_, c_, nullptr) != PFS_OK) return fail(pfs_last_error());
        return commit();
    }

    // The two halves of submit for a producer that forms the bytes itself (pfs_ctx_advect_color_step_rgba8 stores the
    // frame from the kernel that advects the image): begin hands out the device byte buffer of a free slot, commit is
    // called once the producing kernel is enqueued on the legacy default stream.
    unsigned char *begin(const std::string &path)
    {
        std::unique_lock<std::mutex> lk(m_);
        cv_free_.wait(lk, [this] { return !free_.empty(); });
        open_ = free_.front();
        free_.pop_front();
        open_->path = path;
        return open_->dev;
    }

    bool commit()
    {
        Slot *s = open_;
        open_ = nullptr;
        if (!s) return fail("FrameWriter::commit without begin");
        if (cudaEventRecord(s->packed, nullptr) != cudaSuccess ||
            cudaStreamWaitEvent(copy_stream_, s->packed, 0) != cudaSuccess ||
            cudaMemcpyAsync(s->host, s->dev, bytes_, cudaMemcpyDeviceToHost, copy_stream_) != cudaSuccess ||
            cudaEventRecord(s->copied, copy_stream_) != cudaSuccess)
            return fail(cudaGetErrorString(cudaGetLastError()));
        {
            std::lock_guard<std::mutex> lk(m_);
            queued_.push_back(s);
        }
        cv_work_.notify_one();
        return true;
    }

    // Wait until every submitted frame is on disk (or was reported as unwritable) and stop the workers.
    // false only if a CUDA call failed.
    bool finish()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            if (stopping_) return !failed_;
            stopping_ = true;
        }
        cv_work_.notify_all();
        for (auto &t : workers_) t.join();
        return !failed_;
    }

    const std::string &error() const { return error_; }

private:
    struct Slot {
        unsigned char *dev = nullptr, *host = nullptr;
        cudaEvent_t packed = nullptr, copied = nullptr;
        std::string path;
    };

    bool fail(const char *what)
    {
        std::lock_guard<std::mutex> lk(m_);
        failed_ = true;
        error_ = what ? what : "unknown error";
        return false;
    }

    // A frame that cannot be written (missing output directory, full disk) is not fatal: the reference ignores the
    // return value of write_png_from_array (main.cpp:68), finishes the run and exits 0.  Same here, plus a line on stderr.
    void warn_unwritten(const std::string &path)
    {
        std::lock_guard<std::mutex> lk(m_);
        ++unwritten_;
        std::fprintf(stderr, "warning: cannot write %s\n", path.c_str());
    }

    void work()
    {
        for (;;) {
            Slot *s = nullptr;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [this] { return stopping_ || !queued_.empty(); });
                if (queued_.empty()) return;   // stopping and drained
                s = queued_.front();
                queued_.pop_front();
            }
            pngio::png_image header = header_;   // libpng keeps per-write state in the struct: one copy per frame
            if (cudaEventSynchronize(s->copied) != cudaSuccess)
                fail("frame copy failed");
            else if (pngio::write_png_from_bytes(&header, s->path.c_str(), s->host) != 0)
                warn_unwritten(s->path);
            {
                std::lock_guard<std::mutex> lk(m_);
                free_.push_back(s);
            }
            cv_free_.notify_one();
        }
    }

    pngio::png_image header_;
    int w_, h_, c_;
    size_t bytes_;
    bool ok_ = false, failed_ = false, stopping_ = false;
    int unwritten_ = 0;
    std::string error_;
    cudaStream_t copy_stream_ = nullptr;
    std::vector<Slot> slots_;
    Slot *open_ = nullptr;              // slot handed out by begin(), waiting for commit()
    std::deque<Slot *> free_, queued_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_free_, cv_work_;
};
