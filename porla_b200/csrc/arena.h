// Device scratch arena shared by the per-curve MSM drivers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "msm.h"

namespace porla {

// One growable device block, carved per call.  Grown geometrically, never shrunk; sized for
// 180 GB parts (a 2^26-point MSM needs ~5 GB of scratch).
struct Arena {
    uint8_t* base = nullptr;
    size_t cap = 0, used = 0;
    void reset() { used = 0; }
    void reserve(size_t bytes, cudaStream_t stream) {
        if (bytes <= cap) return;
        if (base) {
            PORLA_CUDA(cudaStreamSynchronize(stream));
            PORLA_CUDA(cudaDeviceSynchronize());
            PORLA_CUDA(cudaFree(base));
        }
        size_t want = bytes + bytes / 8 + (1u << 20);
        PORLA_CUDA(cudaMalloc(&base, want));
        cap = want;
    }
    template <class T>
    T* take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        T* p = reinterpret_cast<T*>(base + used);
        used += bytes;
        if (used > cap) {
            fprintf(stderr, "[libmultiexp/porla_b200] FATAL: arena overflow\n");
            abort();
        }
        return p;
    }
    static size_t padded(size_t count, size_t elt) { return (count * elt + 255) & ~(size_t)255; }

    // The block is shared by consecutive calls; a call on ANOTHER stream must not start before the previous user's
    // kernels have finished with it.  acquire() at the start of a call (engine mutex held), release() after its last
    // launch.
    cudaEvent_t done = nullptr;
    cudaStream_t last_stream = nullptr;
    bool in_use = false;
    void acquire(cudaStream_t stream) {
        if (in_use && stream != last_stream) PORLA_CUDA(cudaStreamWaitEvent(stream, done, 0));
    }
    void release(cudaStream_t stream) {
        if (!done) PORLA_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        PORLA_CUDA(cudaEventRecord(done, stream));
        last_stream = stream;
        in_use = true;
    }
};

// Optional per-stage CUDA-event timing of the most recent MSM (bench.py's live roofline figure).
enum { kStageCount = 0, kStageScan, kStageScatter, kStageAccumulate, kStageReduce, kStageFinalize, kNumStages };
struct StageTimer {
    bool enabled = false;
    cudaEvent_t ev[kNumStages + 1] = {};
    bool created = false;
    void mark(int i, cudaStream_t s) {
        if (!enabled) return;
        if (!created) {
            for (auto& e : ev) PORLA_CUDA(cudaEventCreate(&e));
            created = true;
        }
        PORLA_CUDA(cudaEventRecord(ev[i], s));
    }
};

// Per-device engine state.  One process may drive several devices (one worker thread per device inside a call);
// each device has its own scratch arena, its own engine mutex and its own stage events.
struct DeviceCtx {
    std::mutex engine_mu;
    Arena arena;
    StageTimer timer;
};
DeviceCtx& device_ctx();     // of the calling thread's current engine device (device_init / DeviceScope)

}  // namespace porla
