// hostpath.cuh -- host-side interface of the host-resident batch pipeline (see hostpath.cu)
#pragma once
#include "bg_common.cuh"

namespace bg {

struct HostPipe;
int32_t hostpipe_create(HostPipe** out, int32_t device, int32_t H, int64_t chunk_units, int32_t all_rolls, int32_t item_cap, int32_t rows_per_item,
                        int32_t n_streams);
int32_t hostpipe_destroy(HostPipe* p);
int32_t hostpipe_run(HostPipe* p, const int8_t* h_boards, const uint8_t* h_players, const uint8_t* h_rolls, int64_t n_units, const float* prepared,
                     float temperature, uint64_t seed, int32_t* h_actions, int32_t* h_counts, cudaStream_t stream);
int32_t hostpipe_status(HostPipe* p, int32_t* out_status);

}  // namespace bg
