// Contraction kernel instances WITH the masked sub-tile path (triangular rank-reduced blocks, padding of the last column block):
// the same source as contract.cu, compiled as a second translation unit.
#define DLB_CONTRACT_MASKED 1
#include "contract.cu"
