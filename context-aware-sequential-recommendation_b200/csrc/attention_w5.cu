// explicit instantiations of the attention kernels for padded head widths 40, 44, 48
#include "attention.cuh"

namespace cast {
template int dispatch_att<40>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
template int dispatch_att<44>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
template int dispatch_att<48>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
}  // namespace cast
