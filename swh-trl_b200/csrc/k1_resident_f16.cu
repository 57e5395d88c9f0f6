// The fp16 instantiations of the K1 resident kernel (see the note at the top of k1_resident.cu): a second translation
// unit of the same source, so that the two halves of the kernel family compile in parallel.
#define K1_UNIT_F16 1
#include "k1_resident.cu"
