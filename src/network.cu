// network.cu — the translation unit the reference lists for its network (reference src/network.cu is
// 0 bytes, CMakeLists.txt:12-16).  The class is header-only like the reference's tensor; this file
// pins the one instantiation the sm_100a library supports so users link a single copy.
#include<network.cuh>

template class network<float>;
