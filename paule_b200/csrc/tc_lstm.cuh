// Shapes of the H=720 persistent-RNN kernels (forward: tc_lstm_fwd.cu, backward: tc_lstm_bwd.cu).
#pragma once
#include <stdint.h>

namespace paule {
namespace tc {

constexpr int kH = 720;                 // hidden size the tensor-core path is built for (paule/paule.py:124,167)
constexpr int kKPad = 768;              // K padded to 12 k-blocks of 64
constexpr int kNumKB = kKPad / 64;      // 12
constexpr int kRows = 64;               // batch rows per launch (UMMA M)

// forward: 90 CTAs x 8 hidden units, B operand [N=32 (4 gates x 8 units), K=768]
constexpr int kFwdUnits = 8;
constexpr int kFwdCtas = kH / kFwdUnits;            // 90
constexpr int kFwdN = 4 * kFwdUnits;                // 32
constexpr int kFwdSliceBytes = kNumKB * kFwdN * 128;  // 49152

// backward: 23 unit groups x 4 gates = 92 CTAs in clusters of 4 (split-K over the gates),
// B operand [N=32 units, K=768 (the units k of gate g)]
constexpr int kBwdN = 32;
constexpr int kBwdGroups = (kH + kBwdN - 1) / kBwdN;  // 23
constexpr int kBwdCtas = kBwdGroups * 4;              // 92
constexpr int kBwdSliceBytes = kNumKB * kBwdN * 128;  // 49152

// exchange buffer: header + UMMA images [12][64 rows][128 B]
constexpr int kXchgHeader = 4096;   // 12 barrier counters (one 128-byte line per k-block), error flag, trace words
constexpr int kXchgErrOff = 2048;    // int: 0 ok, 1 grid-barrier watchdog fired, 2 mbarrier watchdog fired
constexpr int kXchgTraceOff = 3072;  // PAULE_TC_TRACE builds only
constexpr int kXchgImageBytes = kNumKB * kRows * 128;  // 98304

}  // namespace tc
}  // namespace paule
