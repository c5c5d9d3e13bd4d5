// pf_comm.h -- multi-GPU slab plumbing (NCCL over NVLink): halo rows between
// neighbouring row slabs and all-reduces of the step / CG scalars.  New design
// (the reference has no distributed path at all: SURVEY.md section 8e).
#pragma once

#include "pf_common.cuh"

namespace pf {

struct Comm;

int comm_unique_id(void* out128);
int comm_create(Comm** out, const void* id128, int rank, int nranks, const Slab& s, char* err);
void comm_destroy(Comm* c);
int comm_rank(const Comm* c);
int comm_nranks(const Comm* c);

// Fill my lower halo rows [-from_below, 0) with the top rows of the slab below
// and my upper halo rows [nyl, nyl+from_above) with the bottom rows of the slab
// above.  Collective over neighbours: every rank calls it with the same counts.
int comm_halo(Comm* c, double* field, int from_below, int from_above, cudaStream_t st, char* err);

// In-place all-reduce of n doubles in device memory (sum, or max if is_max).
int comm_allreduce_scalars(Comm* c, double* dev, int n, int is_max, cudaStream_t st, char* err);

}  // namespace pf
