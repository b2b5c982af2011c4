// explicit instantiations of the attention kernels for padded head widths 20, 24, 36
#include "attention.cuh"

namespace cast {
template int dispatch_att<20>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
template int dispatch_att<24>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
template int dispatch_att<36>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
}  // namespace cast
