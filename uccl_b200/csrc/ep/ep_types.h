// Host-safe argument structs of the expert-parallel (EP) kernels.
#pragma once
#include <stdint.h>

#include "../kernels/types.h"

namespace ub {

constexpr int kEpMaxTopk = 32;          // DeepEP HT limit (ep/src/intranode.cu:213)
constexpr int kEpMaxLocalExperts = 1024;  // ep/include/ep_configs.cuh:6
constexpr int kEpMaxBlocks = 256;

enum EpXMode : int {
  EP_X_BF16 = 0,       // bf16 in  -> bf16 out (plain permutation)
  EP_X_FP8_SCALED = 1, // (fp8, float scales[H/128]) in -> same out (DeepEP pre-quantised input)
  EP_X_FUSED_FP8 = 2   // bf16 in -> per-128 amax / scale / e4m3 cast fused into the send -> (fp8, scales) out
};

// One receive arena (a "slot" of the dispatch ring) inside the symmetric heap.
struct EpArena {
  uint64_t x_off;        // [cap][H * elem]
  uint64_t scales_off;   // [cap][H/128] float (fp8 modes)
  uint64_t topk_idx_off; // [cap][K] int64
  uint64_t topk_w_off;   // [cap][K] float
  uint64_t src_idx_off;  // [cap] int32
  int capacity;          // tokens
};

struct EpLayoutArgs {
  const int64_t* topk_idx;  // [T, K]
  int T, K, E, R;
  int32_t* tokens_per_rank;    // [R]
  int32_t* tokens_per_expert;  // [E]
  uint8_t* is_token_in_rank;   // [T, R] (torch.bool)
  int32_t* token_pos;          // [T, R] exclusive position among my tokens going to rank r, -1 if none
  // multi-CTA layout (ep_layout_mc_kernel): device scratch of kEpLayoutScratchWords u32
  //   [0] call epoch, [16 + b] flag of block b, [16 + kEpLayoutMaxBlocks + b * kMaxRanks + r] tokens of block b for rank r
  uint32_t* scratch;
  int tokens_per_block;        // contiguous tokens owned by one CTA (multiple of the block size)
};
constexpr int kEpLayoutMaxBlocks = 64;
constexpr int kEpLayoutThreads = 128;
constexpr int kEpLayoutScratchWords = 16 + kEpLayoutMaxBlocks + kEpLayoutMaxBlocks * kMaxRanks;

// Kernel implementation of the high-throughput dispatch / combine.
enum EpImpl : int {
  EP_IMPL_AUTO = 0,  // pick per (ranks, CTAs): see EpBuffer::pick_impl
  EP_IMPL_REG = 1,   // register path: every warp loads a row with LDG.128 and stores it with STG.128 (ep_kernels.cu)
  EP_IMPL_TMA = 2    // warp-specialised cp.async.bulk pipelines through shared memory (ep_tma_kernels.cu)
};

struct EpDispatchArgs {
  const void* x;
  const float* x_scales;
  const int64_t* topk_idx;    // may be null
  const float* topk_weights;  // may be null
  const int32_t* token_pos;   // [T, R] (non-cached) -- rank-local positions
  int32_t* send_slot;         // [T, R] absolute slot inside each destination arena (written when !cached, read when cached)
  const int32_t* tokens_per_rank;    // [R] device
  const int32_t* tokens_per_expert;  // [E] device (may be null when cached)
  int T, H, K, E;
  int mode;    // EpXMode
  int cached;  // 1: reuse send_slot, skip the count exchange
  EpArena arena;
  uint64_t cnt_tab_off;  // symmetric: [kEpMaxBlocks][R src][R dst] int32
  uint64_t exp_tab_off;  // symmetric: [R src][kEpMaxLocalExperts] int32
  int32_t* rank_prefix;  // local [R][R]: M[src][dst] token counts (handle)
  int32_t* dev_counts;   // local [1 + kEpMaxLocalExperts]: recv total, per-local-expert recv counts
  volatile int32_t* host_counts;  // host-mapped mirror of dev_counts (CPU spins on [0]); may be null
  int expert_alignment;
  int num_worst_tokens;
  int round_scale;  // power-of-two scales (UE8M0-compatible), like DeepEP's round_scale
  int in_stages, out_stages;  // TMA pipeline depth (ep_tma_kernels.cu); 0 = launcher default
};

struct EpCombineArgs {
  uint64_t x_off;            // symmetric: expert outputs [num_recv, H] bf16 (in the combine arena)
  uint64_t topk_w_off;       // symmetric: [num_recv, K] float, or kNoOff
  const int32_t* send_slot;  // [T, R] from dispatch
  const void* bias0;         // optional [T, H] bf16
  const void* bias1;
  void* out;                 // [T, H] bf16
  float* out_topk_w;         // [T, K] or null
  int T, H, K;
  int stages;                // TMA pipeline depth (ep_tma_kernels.cu); 0 = launcher default
};

// ---- low-latency (decode) mode -------------------------------------------------------
struct EpLLDispatchArgs {
  const void* x;             // [T, H] bf16
  const int64_t* topk_idx;   // [T, K]
  int T, H, K, E, M;         // E total experts, M = max dispatch tokens per rank
  int use_fp8, round_scale;
  uint64_t recv_x_off;       // symmetric [E_local][R*M][row_bytes]
  uint64_t recv_scales_off;  // symmetric [E_local][R*M][H/128] float
  uint64_t recv_src_off;     // symmetric [E_local][R*M] int32 (source token index)
  uint64_t cnt_tab_off;      // symmetric [blocks][R][E] int32
  int32_t* send_cnt;         // local [2][E] slot counters (parity double-buffered)
  int parity;
  int64_t* send_pos;         // local [T, K]: (dst rank << 32) | row index at the destination, -1 if unused
  int32_t* recv_count;       // local [E_local]
  int64_t* layout_range;     // local [E_local, R]: (begin << 32) | count
  int phase;                 // EpLLPhase: full kernel, send half, or receive half (return_recv_hook)
  int scale_layout;          // EpLLScaleLayout of recv_scales
  long long* wait_stats;     // optional [R] int64: cycles spent waiting for each source rank (DeepEP dispatch_wait_recv_cost_stats)
};

// Hook split of the low-latency kernels (reference: LOW_LATENCY_SEND_PHASE / _RECV_PHASE, ep/src/internode_ll.cu:115,458-464)
enum EpLLPhase : int { EP_LL_FULL = 0, EP_LL_SEND = 1, EP_LL_RECV = 2 };
// recv_scales formats.  DeepEP hands out the column-major ("transposed", TMA-friendly for the grouped GEMM) forms:
//   ROW_MAJOR  float  [E_local][R*M][H/128]
//   COL_MAJOR  float  [E_local][H/128][R*M]            (viewed as [E_local][R*M][H/128] with strides (.., 1, R*M))
//   COL_UE8M0  int32  [E_local][H/512][R*M], each word packs the exponent bytes of 4 consecutive 128-channel groups
enum EpLLScaleLayout : int { EP_LL_SCALES_ROW_MAJOR = 0, EP_LL_SCALES_COL_MAJOR = 1, EP_LL_SCALES_COL_UE8M0 = 2 };

struct EpLLCombineArgs {
  uint64_t x_off;             // symmetric [E_local][R*M][H] bf16 expert outputs
  const int64_t* send_pos;    // [T, K] from dispatch
  const float* topk_weights;  // [T, K]
  void* out;                  // [T, H] bf16
  int T, H, K;
  int phase;                  // EpLLPhase
  long long* wait_stats;      // optional [R] int64 (DeepEP combine_wait_recv_cost_stats)
};

// Packs a user tensor of expert outputs into the symmetric combine buffer: only the rows that hold tokens
// (layout_range) are copied, not the worst-case [E_local][R*M] block.
struct EpLLPackArgs {
  const void* src;              // [E_local][R*M][H] bf16, any device memory
  void* dst;                    // same shape inside the symmetric heap
  const int64_t* layout_range;  // [E_local, R]
  int E_local, R, M, H;
  int logfmt;                   // 1: apply DeepEP's simulated LogFMT-10 cast to the rows on the way (src may equal dst)
};

}  // namespace ub
