// G2 instantiation of the curve kernels (coordinates in Fp2).
#include "group_impl.cuh"
namespace bmpc {
template struct GroupOps<Fp2>;
}
