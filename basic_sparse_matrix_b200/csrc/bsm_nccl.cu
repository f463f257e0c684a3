// bsm_nccl.cu — the one optional collective of the path: all-gather(v) of C row blocks over
// NVLink 5 / NVSwitch (north_star: "used only when the caller asks for a gathered result").
// One process per GPU; the communicator is bootstrapped from a unique id the caller ships to the
// other ranks (torch.distributed in bench.py / tests, any transport for a Rust host).
// nnz-balanced partitions have unequal row counts, so the gather is a group of ncclBroadcast
// calls, one per root, each landing directly in its slot of the row-major result.
#include <nccl.h>

#include <cstring>
#include <string>

#include "bsm_common.cuh"

struct bsm_comm {
    ncclComm_t comm = nullptr;
    int nranks = 0, rank = 0;
    int *token = nullptr;   // 4 bytes of HBM for the stream-ordered barrier
};

using namespace bsm;

#define BSM_NCCL(expr)                                                                            \
    do {                                                                                          \
        ncclResult_t _r = (expr);                                                                 \
        if (_r != ncclSuccess)                                                                    \
            return ::bsm::fail(BSM_ERR_NCCL, std::string(#expr) + ": " + ncclGetErrorString(_r)); \
    } while (0)

extern "C" {

int bsm_comm_unique_id(char id[128])
{
    static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is expected to be 128 bytes");
    if (!id) return fail(BSM_ERR_INVALID_ARGUMENT, "comm_unique_id: null");
    ncclUniqueId uid;
    BSM_NCCL(ncclGetUniqueId(&uid));
    memcpy(id, &uid, sizeof(uid));
    return BSM_OK;
}

int bsm_comm_init(const char id[128], int nranks, int rank, bsm_comm **out)
{
    BSM_TRY(ensure_init());
    if (!id || !out || nranks < 1 || rank < 0 || rank >= nranks) return fail(BSM_ERR_INVALID_ARGUMENT, "comm_init: bad arguments");
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    bsm_comm *c = new bsm_comm();
    c->nranks = nranks;
    c->rank = rank;
    ncclResult_t r = ncclCommInitRank(&c->comm, nranks, uid, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail(BSM_ERR_NCCL, std::string("ncclCommInitRank: ") + ncclGetErrorString(r));
    }
    if (cudaMalloc(&c->token, 4) != cudaSuccess || cudaMemset(c->token, 0, 4) != cudaSuccess) {
        ncclCommDestroy(c->comm);
        delete c;
        return fail(BSM_ERR_CUDA, "comm_init: cannot allocate the barrier token");
    }
    *out = c;
    return BSM_OK;
}

int bsm_comm_free(bsm_comm *c)
{
    if (!c) return BSM_OK;
    if (c->token) cudaFree(c->token);
    if (c->comm) ncclCommDestroy(c->comm);
    delete c;
    return BSM_OK;
}

int bsm_allgather_rows(bsm_comm *c, const bsm_dense *local_block, const uint64_t *bounds, bsm_dense *full)
{
    BSM_TRY(ensure_init());
    if (!c || !local_block || !bounds || !full) return fail(BSM_ERR_INVALID_ARGUMENT, "allgather_rows: null argument");
    if (local_block->dtype != full->dtype) return fail(BSM_ERR_DTYPE_MISMATCH, "allgather_rows: dtype mismatch");
    const uint64_t my_rows = bounds[c->rank + 1] - bounds[c->rank];
    if (local_block->rows != my_rows || local_block->cols != full->cols || full->rows != bounds[c->nranks] ||
        local_block->ld != full->ld)
        return fail(BSM_ERR_INCORRECT_DIMENSIONS, "allgather_rows: block / result dimensions do not match the partition");
    const size_t s = dtype_size(full->dtype);
    const size_t row_bytes = (size_t)full->ld * s;
    cudaStream_t sm = rt().stream;
    // own block into its slot, then every root broadcasts its slot in place
    if (my_rows)
        BSM_CUDA(cudaMemcpyAsync((char *)full->data + bounds[c->rank] * row_bytes, local_block->data, my_rows * row_bytes,
                                 cudaMemcpyDeviceToDevice, sm));
    BSM_NCCL(ncclGroupStart());
    ncclResult_t res = ncclSuccess;
    for (int root = 0; root < c->nranks && res == ncclSuccess; ++root) {
        const uint64_t r0 = bounds[root], r1 = bounds[root + 1];
        if (r1 == r0) continue;
        char *slot = (char *)full->data + r0 * row_bytes;
        res = ncclBroadcast(slot, slot, (r1 - r0) * row_bytes, ncclChar, root, c->comm, sm);
    }
    const ncclResult_t end = ncclGroupEnd();   // always close the group, also after a failed broadcast
    if (res != ncclSuccess) return fail(BSM_ERR_NCCL, std::string("ncclBroadcast: ") + ncclGetErrorString(res));
    if (end != ncclSuccess) return fail(BSM_ERR_NCCL, std::string("ncclGroupEnd: ") + ncclGetErrorString(end));
    return BSM_OK;
}

// Stream-ordered barrier across the ranks (a 4-byte all-reduce): after it, everything every rank
// enqueued before its own call — e.g. the P2P stores of bsm_spmm_scatter — has completed.
int bsm_comm_barrier(bsm_comm *c)
{
    BSM_TRY(ensure_init());
    if (!c) return fail(BSM_ERR_INVALID_ARGUMENT, "comm_barrier: null communicator");
    BSM_NCCL(ncclAllReduce(c->token, c->token, 1, ncclInt, ncclSum, c->comm, rt().stream));
    return BSM_OK;
}

}  // extern "C"
