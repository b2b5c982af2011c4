// explicit instantiations of the attention kernels for padded head widths 128, 256
#include "attention.cuh"

namespace cast {
template int dispatch_att<128>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
template int dispatch_att<256>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
}  // namespace cast
