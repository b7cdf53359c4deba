// Device program ISA: what the tape compiler (tape_compile.cpp) emits and the interpreter
// kernel (interp.cu) executes.  One 64-bit word per instruction (CONST carries a second word
// with the IEEE-754 bits of the literal).
//
//   bits  0.. 7  opcode (enum DevOp)
//   bits  8..25  D  (18 bits)  destination shared slot | output index (OUTPUT) | D_NONE
//   bits 26..44  A  (19 bits)  source shared slot | input index (INPUT) | scratch slot (FILL/SPILL) | F_ACC
//   bits 45..63  B  (19 bits)  source shared slot | nonzero index (INPUT/OUTPUT)       | F_ACC
//
// All arithmetic instructions read and write the per-instance work vector in SHARED memory; the
// allocator inserts FILL/SPILL instructions that move values between shared slots and the
// per-instance global scratch, and re-materialises constants and inputs instead of spilling them.
// F_ACC as a source means "the result of the previous instruction" (kept in a register);
// D_NONE as destination means "only the next instruction reads it" (never stored).
#pragma once
#include <cstdint>

namespace ccu {

enum DevOp : uint8_t {
  D_END = 0,
  D_CONST = 1,   // D <- literal (next word)
  D_INPUT = 2,   // D <- arg[A][B]   (0 when arg[A] is NULL)
  D_OUTPUT = 3,  // res[D][B] <- A   (skipped when res[D] is NULL)
  D_FILL = 4,    // D <- scratch[A]
  D_SPILL = 5,   // scratch[A(=field A)] <- B(slot or ACC)
  // binary: D <- A op B
  D_BIN_FIRST = 16,
  D_ADD = 16, D_SUB, D_MUL, D_DIV, D_POW, D_LT, D_LE, D_EQ, D_NE, D_AND, D_OR, D_FMOD, D_COPYSIGN,
  D_IF_ELSE_ZERO, D_FMIN, D_FMAX, D_ATAN2, D_HYPOT, D_REMAINDER,
  D_BIN_LAST = D_REMAINDER,
  // unary: D <- op A
  D_UN_FIRST = 64,
  D_COPY = 64, D_NEG, D_EXP, D_LOG, D_SQRT, D_SQ, D_TWICE, D_SIN, D_COS, D_TAN, D_ASIN, D_ACOS, D_ATAN,
  D_NOT, D_FLOOR, D_CEIL, D_FABS, D_SIGN, D_ERF, D_INV, D_SINH, D_COSH, D_TANH, D_ASINH, D_ACOSH, D_ATANH,
  D_ERFINV, D_LOG1P, D_EXPM1,
  D_UN_LAST = D_EXPM1,
};

constexpr int kDBits = 18, kABits = 19, kBBits = 19;
constexpr uint32_t D_NONE = (1u << kDBits) - 1;  // destination: do not store
constexpr uint32_t F_ACC = (1u << kABits) - 1;   // source: previous result
constexpr uint32_t kMaxSharedSlots = D_NONE - 1;
constexpr uint32_t kMaxFieldIndex = F_ACC - 1;   // input/output/nonzero/scratch indices

static inline uint64_t enc(uint32_t op, uint32_t d, uint32_t a, uint32_t b) {
  return (uint64_t)op | ((uint64_t)d << 8) | ((uint64_t)a << 26) | ((uint64_t)b << 45);
}
#define CCU_DEC_OP(w) ((uint32_t)((w) & 0xffu))
#define CCU_DEC_D(w) ((uint32_t)(((w) >> 8) & 0x3ffffu))
#define CCU_DEC_A(w) ((uint32_t)(((w) >> 26) & 0x7ffffu))
#define CCU_DEC_B(w) ((uint32_t)((w) >> 45))

}  // namespace ccu
