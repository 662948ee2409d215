// learner.cuh -- host-side interface of the TD(0) learner (see learner.cu)
#pragma once
#include "bg_common.cuh"

namespace bg {

struct Learner;
int learner_device(const Learner* L);  // the device the handle lives on

int32_t learner_create(Learner** out, int32_t device, int32_t H, float lr, float gamma, float grad_clip);
int32_t learner_destroy(Learner* L);
int32_t learner_set_parameters(Learner* L, const float* packed_dev, int32_t reset_optimizer, cudaStream_t s);
int32_t learner_get_parameters(Learner* L, float* packed_dev, cudaStream_t s);
int32_t learner_get_optimizer(Learner* L, float* m_dev, float* v_dev, int64_t* step_dev, cudaStream_t s);
int32_t learner_update(Learner* L, const int8_t* boards, const uint8_t* flags_or_meta, const float* reward, const int64_t* ep_offsets,
                       const int32_t* ep_len, int64_t n_eps, int32_t records, float* out_metrics, int32_t* out_status, cudaStream_t s);

}  // namespace bg
