// G2 instantiation of the curve kernels (coordinates in Fp2).
// Every Fp2 product of this unit calls the out-of-line Fp product three times (field.cuh:
// BMPC_FP2_CALLS) instead of inlining three bodies: the batched-affine kernel shrinks from 197 KB to a
// size the instruction cache holds (`no_instruction` was its top stall), and since the routine takes its
// operands in registers the calls cost ~10 % more instructions, not the stack traffic they used to.
#define BMPC_FP2_CALLS 1
#include "group_impl.cuh"
namespace bmpc {
template struct GroupOps<Fp2>;
}
