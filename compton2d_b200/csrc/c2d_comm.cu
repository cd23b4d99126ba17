// c2d_comm.cu -- the collectives of the hot path, inside the library (host code).
//
// What they replace in the reference (file:line relative to /root/reference/src_20121026/):
//   c2d_tally_reduce      cens_add_up (update2d.f:1931-1999: edep, ecens, n_field), E_add_up (update2d.f:2002-2083:
//                         erlk*, E_IC, ...), xec_add (xec2d.f:325-355: Ed_ref, Ed_in, ecens, npcen) and graphics_collect
//                         (xec2d.f:371-394: edout, fout) -- ~20 blocking MPI_REDUCE calls per step, several of them
//                         element by element -- as ONE ncclAllReduce of the packed FP64 tally buffer and one of the
//                         packed int64 buffer, in place in HBM on the library's stream.
//   comm_allgather        FP_recv_result + FP_end_bcast (fp_mpi.f:926,1017-1079): the updated electron state of the
//                         zones every rank solved, gathered to all ranks.
//   3rd-split rebalancing no counterpart (the reference's farm hands zones to whoever is idle, MPI_ANY_SOURCE): the
//                         3rd-split tasks of all ranks are all-gathered and dealt out again sub-child by sub-child.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already holds -- torch's in the Python host -- or the
// system one), so the library still loads on a box without NCCL and a single-GPU run never touches it.  A host that owns
// another transport (the reference's MPI, gloo in the CPU tests of the host logic) plugs it in with c2d_comm_init_ext.
#include "c2d_ctx.h"

#include <cstdio>
#include <cstring>

#ifndef C2D_EMU
#include <dlfcn.h>
#include <nccl.h>
#endif

namespace c2d {

struct Comm {
    int rank = 0, world = 1;
    bool ext = false;
    c2d_comm_ext vt{};
#ifndef C2D_EMU
    void *dl = nullptr;
    ncclComm_t comm = nullptr;
    ncclResult_t (*pCommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*pCommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*pAllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*pAllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*pGetErrorString)(ncclResult_t) = nullptr;
#endif
};

#ifndef C2D_EMU
static void *nccl_open()
{
    // the copy the process already mapped (torch ships its own libnccl.so.2), else whatever the loader finds
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    return h;
}
#define NCCLCHK(c, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) \
    return fail(C2D_ECOMM, "%s failed: %s", #call, (c)->pGetErrorString ? (c)->pGetErrorString(r_) : "?"); } while (0)
#endif

int comm_world(const Ctx *c) { return c->comm ? c->comm->world : 1; }
int comm_rank(const Ctx *c) { return c->comm ? c->comm->rank : 0; }

void comm_free(Ctx *c)
{
    if (!c->comm) return;
#ifndef C2D_EMU
    if (c->comm->comm && c->comm->pCommDestroy) c->comm->pCommDestroy(c->comm->comm);
#endif
    delete c->comm;
    c->comm = nullptr;
}

// kind: 0 = FP64 sum, 1 = int64 sum, 2 = FP64 max.  send == recv allowed (in place).
int comm_allreduce(Ctx *c, const void *send, void *recv, size_t n, int kind)
{
    Comm *m = c->comm;
    if (!m || m->world == 1) {
        if (send != recv) {
            cudaError_t e = cudaMemcpyAsync(recv, send, n * 8, cudaMemcpyDeviceToDevice, c->stream);
            if (e != cudaSuccess) return fail(C2D_ECUDA, "cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
        }
        return 0;
    }
    if (m->ext) {
        if (!m->vt.allreduce) return fail(C2D_ECOMM, "external transport has no allreduce");
        const int rc = m->vt.allreduce(m->vt.user, send, recv, (int64_t)n, kind, (void *)c->stream);
        return rc ? fail(C2D_ECOMM, "external allreduce failed (%d)", rc) : 0;
    }
#ifndef C2D_EMU
    NCCLCHK(m, m->pAllReduce(send, recv, n, kind == 1 ? ncclInt64 : ncclFloat64, kind == 2 ? ncclMax : ncclSum, m->comm, c->stream));
    return 0;
#else
    return fail(C2D_ECOMM, "no transport");
#endif
}

// recv = [world][bytes] (rank order)
int comm_allgather(Ctx *c, const void *send, void *recv, size_t bytes)
{
    Comm *m = c->comm;
    if (!m || m->world == 1) {
        if (send != recv && bytes) {
            cudaError_t e = cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, c->stream);
            if (e != cudaSuccess) return fail(C2D_ECUDA, "cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
        }
        return 0;
    }
    if (bytes == 0) return 0;
    if (m->ext) {
        if (!m->vt.allgather) return fail(C2D_ECOMM, "external transport has no allgather");
        const int rc = m->vt.allgather(m->vt.user, send, recv, (int64_t)bytes, (void *)c->stream);
        return rc ? fail(C2D_ECOMM, "external allgather failed (%d)", rc) : 0;
    }
#ifndef C2D_EMU
    NCCLCHK(m, m->pAllGather(send, recv, bytes, ncclUint8, m->comm, c->stream));
    return 0;
#else
    return fail(C2D_ECOMM, "no transport");
#endif
}

}  // namespace c2d

using namespace c2d;

extern "C" {

int c2d_comm_unique_id(void *id128)
{
    if (!id128) return fail(C2D_EINVAL, "NULL id buffer");
#ifndef C2D_EMU
    static_assert(sizeof(ncclUniqueId) == C2D_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    void *h = nccl_open();
    if (!h) return fail(C2D_ECOMM, "libnccl.so.2 not found: %s", dlerror());
    auto get = reinterpret_cast<ncclResult_t (*)(ncclUniqueId *)>(dlsym(h, "ncclGetUniqueId"));
    if (!get) return fail(C2D_ECOMM, "ncclGetUniqueId not found");
    ncclUniqueId id;
    const ncclResult_t r = get(&id);
    if (r != ncclSuccess) return fail(C2D_ECOMM, "ncclGetUniqueId failed (%d)", (int)r);
    memcpy(id128, &id, sizeof id);
    return 0;
#else
    memset(id128, 0, C2D_COMM_ID_BYTES);
    return fail(C2D_ECOMM, "no NCCL in this build");
#endif
}

int c2d_comm_init(c2d_ctx *h, const void *id128, int rank, int world)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c) return fail(C2D_EINVAL, "NULL ctx");
    if (world < 1 || rank < 0 || rank >= world) return fail(C2D_EINVAL, "rank %d of %d", rank, world);
    comm_free(c);
    Comm *m = new Comm();
    m->rank = rank; m->world = world;
    c->comm = m;
    if (world == 1) return 0;
#ifndef C2D_EMU
    if (!id128) return fail(C2D_EINVAL, "NULL id");
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(C2D_ECUDA, "cudaSetDevice(%d) failed", c->device);
    m->dl = nccl_open();
    if (!m->dl) { comm_free(c); return fail(C2D_ECOMM, "libnccl.so.2 not found: %s", dlerror()); }
    m->pCommInitRank = reinterpret_cast<decltype(m->pCommInitRank)>(dlsym(m->dl, "ncclCommInitRank"));
    m->pCommDestroy = reinterpret_cast<decltype(m->pCommDestroy)>(dlsym(m->dl, "ncclCommDestroy"));
    m->pAllReduce = reinterpret_cast<decltype(m->pAllReduce)>(dlsym(m->dl, "ncclAllReduce"));
    m->pAllGather = reinterpret_cast<decltype(m->pAllGather)>(dlsym(m->dl, "ncclAllGather"));
    m->pGetErrorString = reinterpret_cast<decltype(m->pGetErrorString)>(dlsym(m->dl, "ncclGetErrorString"));
    if (!m->pCommInitRank || !m->pCommDestroy || !m->pAllReduce || !m->pAllGather) {
        comm_free(c);
        return fail(C2D_ECOMM, "NCCL entry points missing");
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    const ncclResult_t r = m->pCommInitRank(&m->comm, world, id, rank);
    if (r != ncclSuccess) {
        const char *s = m->pGetErrorString ? m->pGetErrorString(r) : "?";
        comm_free(c);
        return fail(C2D_ECOMM, "ncclCommInitRank(rank %d of %d) failed: %s", rank, world, s);
    }
    return 0;
#else
    (void)id128;
    comm_free(c);
    return fail(C2D_ECOMM, "no NCCL in this build: use c2d_comm_init_ext");
#endif
}

int c2d_comm_init_ext(c2d_ctx *h, int rank, int world, const c2d_comm_ext *vt)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !vt) return fail(C2D_EINVAL, "NULL argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(C2D_EINVAL, "rank %d of %d", rank, world);
    comm_free(c);
    Comm *m = new Comm();
    m->rank = rank; m->world = world; m->ext = true; m->vt = *vt;
    c->comm = m;
    return 0;
}

int c2d_comm_info(c2d_ctx *h, int *rank, int *world)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c) return fail(C2D_EINVAL, "NULL ctx");
    if (rank) *rank = comm_rank(c);
    if (world) *world = comm_world(c);
    return 0;
}

int c2d_comm_destroy(c2d_ctx *h)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c) return 0;
    comm_free(c);
    return 0;
}

}  // extern "C"
