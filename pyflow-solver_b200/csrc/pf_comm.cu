// pf_comm.cu -- NCCL plumbing for row-slab decomposition (see pf_comm.h).
#include "pf_comm.h"

#include <nccl.h>
#include <stdio.h>
#include <string.h>

namespace pf {

struct Comm {
  ncclComm_t nccl = nullptr;
  int rank = 0, nranks = 1;
  Slab s{};
};

static_assert(sizeof(ncclUniqueId) == 128, "pf_comm_unique_id hands out 128 bytes");

#define PF_NCCL(err, call)                                                          \
  do {                                                                              \
    ncclResult_t r__ = (call);                                                      \
    if (r__ != ncclSuccess) {                                                       \
      if (err) snprintf((err), 512, "%s: %s", #call, ncclGetErrorString(r__));      \
      return PF_ERR_NCCL;                                                           \
    }                                                                               \
  } while (0)

int comm_unique_id(void* out128) {
  if (!out128) return PF_ERR_ARG;
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return PF_ERR_NCCL;
  memcpy(out128, &id, sizeof(id));
  return PF_OK;
}

int comm_create(Comm** out, const void* id128, int rank, int nranks, const Slab& s, char* err) {
  if (s.nyl < HALO) {
    if (err) snprintf(err, 512, "slab of %d rows is thinner than the %d-row halo", s.nyl, HALO);
    return PF_ERR_ARG;
  }
  Comm* c = new Comm();
  c->rank = rank;
  c->nranks = nranks;
  c->s = s;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclResult_t r = ncclCommInitRank(&c->nccl, nranks, id, rank);
  if (r != ncclSuccess) {
    if (err) snprintf(err, 512, "ncclCommInitRank: %s", ncclGetErrorString(r));
    delete c;
    return PF_ERR_NCCL;
  }
  *out = c;
  return PF_OK;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  if (c->nccl) ncclCommDestroy(c->nccl);
  delete c;
}

int comm_rank(const Comm* c) { return c->rank; }
int comm_nranks(const Comm* c) { return c->nranks; }

int comm_halo(Comm* c, double* f, int from_below, int from_above, cudaStream_t st, char* err) {
  if (from_below > HALO || from_above > HALO || from_below < 0 || from_above < 0) return PF_ERR_ARG;
  const Slab& s = c->s;
  const bool has_below = c->rank > 0, has_above = c->rank < c->nranks - 1;
  const size_t P = (size_t)s.pitch;
  PF_NCCL(err, ncclGroupStart());
  if (has_above) {
    if (from_below > 0)   // my top rows become the lower halo of the slab above
      PF_NCCL(err, ncclSend(f + cell(s, s.nyl - from_below, 0), from_below * P, ncclDouble, c->rank + 1, c->nccl, st));
    if (from_above > 0)
      PF_NCCL(err, ncclRecv(f + cell(s, s.nyl, 0), from_above * P, ncclDouble, c->rank + 1, c->nccl, st));
  }
  if (has_below) {
    if (from_above > 0)   // my bottom rows become the upper halo of the slab below
      PF_NCCL(err, ncclSend(f + cell(s, 0, 0), from_above * P, ncclDouble, c->rank - 1, c->nccl, st));
    if (from_below > 0)
      PF_NCCL(err, ncclRecv(f + cell(s, -from_below, 0), from_below * P, ncclDouble, c->rank - 1, c->nccl, st));
  }
  PF_NCCL(err, ncclGroupEnd());
  return PF_OK;
}

int comm_allreduce_scalars(Comm* c, double* dev, int n, int is_max, cudaStream_t st, char* err) {
  PF_NCCL(err, ncclAllReduce(dev, dev, (size_t)n, ncclDouble, is_max ? ncclMax : ncclSum, c->nccl, st));
  return PF_OK;
}

}  // namespace pf
