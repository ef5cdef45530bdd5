// Planner / validator / bounds checker / simulator of csrc/ukernel/uk_plan.cc as a plain C++ program, so that it can run
// under ASan + UBSan without Python or CUDA (the reference's unit layer: experimental/ukernel/src/ccl/test/unit/
// test_components.cc:240-635 -- planner, lowering, ring-allreduce simulator).
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -Iuccl_b200/csrc -I/usr/local/cuda/include \
//       tests/cpp/uk_plan_test.cc uccl_b200/csrc/ukernel/uk_plan.cc uccl_b200/csrc/coll/host_coll.cc ... -o t && ./t
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels/types.h"
#include "ukernel/uk_plan.h"

using namespace ub;

static int failures = 0;
#define EXPECT(c)                                                  \
  do {                                                             \
    if (!(c)) {                                                    \
      std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c);     \
      ++failures;                                                  \
    }                                                              \
  } while (0)

static void run_allreduce(int n, int lanes, UkAlgo algo, size_t count) {
  std::vector<UkPlan> plans;
  for (int r = 0; r < n; ++r) {
    UkPlanParams p;
    p.nranks = n, p.rank = r, p.nlanes = lanes, p.tile_bytes = 4096, p.elem_size = 4, p.algo = algo;
    plans.push_back(uk_plan_allreduce(count * 4, p));
  }
  EXPECT(uk_validate(plans).empty());
  std::vector<std::vector<float>> in(n, std::vector<float>(count)), out(n, std::vector<float>(count, -1.f));
  std::vector<std::vector<char>> sc(n);
  UkSimBuffers b;
  for (int r = 0; r < n; ++r) {
    for (size_t i = 0; i < count; ++i) in[r][i] = (float)(r + 1) * (float)(i % 13);
    sc[r].assign(plans[r].scratch_bytes + 16, 0);
    b.in.push_back((char*)in[r].data());
    b.out.push_back((char*)out[r].data());
    b.scratch.push_back(sc[r].data());
    EXPECT(uk_check_bounds(plans[r], count * 4, count * 4, plans[r].scratch_bytes, lanes, 4).empty());
  }
  EXPECT(uk_simulate(plans, b, kF32, kSum).empty());
  const float tot = (float)(n * (n + 1) / 2);
  for (int r = 0; r < n; ++r)
    for (size_t i = 0; i < count; ++i)
      if (std::fabs(out[r][i] - tot * (float)(i % 13)) > 1e-3f) {
        EXPECT(!"allreduce result");
        return;
      }
}

int main() {
  for (int n : {2, 3, 4, 8})
    for (int lanes : {1, 3})
      for (UkAlgo a : {UkAlgo::Ring, UkAlgo::FullMesh}) run_allreduce(n, lanes, a, 3000 + 16 * n);

  // bounds checker on hand-made plans
  UkPlan p;
  p.nranks = 2, p.rank = 0, p.nlanes = 1, p.scratch_bytes = 64;
  UkPlanOp c;
  c.kind = UkPlanOp::Copy, c.bytes = 32, c.dst = {UkBuf::Out, 0}, c.src = {UkBuf::In, 32};
  p.ops = {c};
  EXPECT(uk_check_bounds(p, 64, 64, 64, 1, 1).empty());
  p.ops[0].src.off = 48;  // 48 + 32 > 64
  EXPECT(!uk_check_bounds(p, 64, 64, 64, 1, 1).empty());
  p.ops[0].src.off = 8;   // misaligned
  EXPECT(!uk_check_bounds(p, 64, 64, 64, 1, 1).empty());
  p.ops[0].src.off = ~0ull - 15;  // offset arithmetic must not wrap
  EXPECT(!uk_check_bounds(p, 64, 64, 64, 1, 1).empty());
  p.ops[0].src.off = 0;
  p.ops[0].lane = 1;
  EXPECT(!uk_check_bounds(p, 64, 64, 64, 1, 1).empty());
  p.ops[0].lane = 0;
  EXPECT(!uk_check_bounds(p, 64, 64, 32, 1, 1).empty());  // plan wants more scratch than the communicator has
  UkPlanOp s;
  s.kind = UkPlanOp::Send, s.peer = 0, s.bytes = 16, s.dst = {UkBuf::Scratch, 0}, s.src = {UkBuf::In, 0};
  p.ops = {s};
  EXPECT(!uk_check_bounds(p, 64, 64, 64, 1, 1).empty());  // send to self
  p.ops[0].peer = 1;
  EXPECT(uk_check_bounds(p, 64, 64, 64, 1, 1).empty());
  UkPlanOp rd;
  rd.kind = UkPlanOp::Reduce, rd.bytes = 6, rd.dst = {UkBuf::Out, 0}, rd.src = {UkBuf::In, 0}, rd.src2 = {UkBuf::In, 16};
  p.ops = {rd};
  EXPECT(!uk_check_bounds(p, 64, 64, 64, 1, 4).empty());  // 6 bytes is not a multiple of 4-byte elements

  // unmatched send is caught across ranks, and a deadlock by the simulator
  UkPlan a0, a1;
  a0.nranks = a1.nranks = 2, a0.rank = 0, a1.rank = 1, a0.nlanes = a1.nlanes = 1, a0.scratch_bytes = a1.scratch_bytes = 64;
  a0.ops = {s};
  a0.ops[0].peer = 1;
  EXPECT(!uk_validate({a0, a1}).empty());
  UkPlanOp r0, r1;
  r0.kind = r1.kind = UkPlanOp::Recv, r0.peer = 1, r1.peer = 0, r0.bytes = r1.bytes = 16;
  UkPlanOp s1 = s;
  s1.peer = 0;
  a0.ops = {r0, a0.ops[0]};  // both ranks wait before they send
  a1.ops = {r1, s1};
  EXPECT(uk_validate({a0, a1}).empty());
  std::vector<char> m0(64), m1(64), q0(80), q1(80);
  UkSimBuffers sb;
  sb.in = {m0.data(), m1.data()}, sb.out = {m0.data(), m1.data()}, sb.scratch = {q0.data(), q1.data()};
  EXPECT(uk_simulate({a0, a1}, sb, kU8, kSum).find("deadlock") != std::string::npos);

  std::printf(failures ? "FAILED (%d)\n" : "PASS\n", failures);
  return failures ? 1 : 0;
}
