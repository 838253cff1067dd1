// G1 instantiation of the curve kernels (coordinates in Fp).
#include "group_impl.cuh"
namespace bmpc {
template struct GroupOps<Fp>;
}
