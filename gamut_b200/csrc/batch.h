// batch.h -- result object of the batched device-resident decoders (opaque `gb200_batch` in the C ABI).
#pragma once
#include "common.h"
#include "../../include/gamut_b200.h"
#include <vector>

struct gb200_batch {
    std::vector<gb200_image_desc> images;
    std::vector<void*> device_allocs;   // owned device memory (output arena last)
    cudaStream_t stream = nullptr;
    double host_parse_ms = 0, device_ms = 0;
    float phase_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // per-phase device time (CUDA events), format specific
    ~gb200_batch() { for (void* p : device_allocs) gb::dev_free(p); }
};
