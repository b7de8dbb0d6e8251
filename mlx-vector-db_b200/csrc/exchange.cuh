// Candidate exchange over NVLink peer memory: entry points shared with search.cu (vs_exchange_result).
#pragma once
#include "common.cuh"

extern "C" {
int vs_exchange_push(int device, const void* src, int64_t bytes, void* const* peer_dst, void* const* peer_flag,
                     int G, uint32_t step, void* counter, void* stream);
int vs_exchange_wait(int device, const void* flags, int G, uint32_t step, void* stream);
int vs_exchange_wait_merge(int device, int metric, const void* flags, int G, uint32_t step, const void* blocks,
                           int64_t block_words, int B, int k, float* out_scores, int32_t* out_ids, void* stream);
}
