// Launchers shared between aux.cu, framegen.cu and pipeline.cu.
#pragma once
#include "common.cuh"

namespace ldpcb {

int launch_tally_nms(ldpcb_handle* h, const uint32_t* bits, const uint8_t* syn, const uint8_t* iters,
                     const uint32_t* truth, int64_t B, uint64_t* counters, cudaStream_t st);
int launch_tally_final(ldpcb_handle* h, const uint32_t* bits, const uint8_t* syn, const int32_t* best_tep,
                       int osd_order, int tep_order, const uint32_t* truth, int64_t B, uint64_t* counters,
                       cudaStream_t st);
size_t select_temp_bytes(int64_t B);
int launch_select(ldpcb_handle* h, const uint8_t* flags, int64_t B, int32_t* idx, int32_t* count, void* temp,
                  cudaStream_t st);
int launch_gen(ldpcb_handle* h, uint64_t seed, uint64_t first_frame, int64_t B, float ebn0_db, float* llr,
               uint32_t* cw_bits, cudaStream_t st);

}  // namespace ldpcb
