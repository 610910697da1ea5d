// cmgpu_comm.cu -- the one collective of the path: per-stream meter results gathered to one rank over
// NCCL (NVLink 5 / NVSwitch), from the C layer.
//
// Streams are independent (transform.c:36-52, vumeter.c:35-57 keep all state per object), so ranks own
// contiguous stream ranges and the data path exchanges nothing. Once per reporting interval every rank
//   1. takes the raw meter rows of its active streams on its compute stream (copy + reset, one kernel,
//      cmgpu_post.cu) -- ordered after every tick queued so far, with no host round trip;
//   2. sends them to the root (grouped ncclSend / ncclRecv on the same stream: 48 B per stereo stream,
//      393 KB per rank at 65,536 streams over 8 GPUs -- latency-, not bandwidth-bound);
// and the root copies the gathered table to pinned host memory once, decodes the position keys and
// finalises dB with the reference's expression and libm: what arrives is, per stream, what
// coolmic_vumeter_result() fills (vumeter.h:48-83).
#include "cmgpu_ctx.h"

#include <nccl.h>

#include <cerrno>
#include <new>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unistd.h>
#include <sys/stat.h>

using cmgpu::fail;

struct cmgpu_comm {
    ncclComm_t nccl = nullptr;
    bool owned = true;
    int device = 0, rank = 0, size = 1;
    cudaStream_t st = nullptr;            // plumbing collectives (barrier, max); the gather runs on the context's stream
    double *d_scalars = nullptr;          // [64] all-reduce scratch
    unsigned *d_counts = nullptr;         // [size] streams per rank
    unsigned *h_counts = nullptr;         // pinned
    unsigned long long *d_gather = nullptr;   // root: rows of every rank
    uint64_t *h_gather = nullptr;         // root: pinned mirror
    size_t gather_u64 = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    std::mutex mu;
};

#define NC(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r__ = (call);                                                                 \
        if (r__ != ncclSuccess)                                                                    \
            return fail(CMGPU_ERR_GENERIC, "%s failed: %s", #call, ncclGetErrorString(r__));       \
    } while (0)

namespace cmgpu {
static __global__ void set_u32(unsigned *p, unsigned v) { *p = v; }
}  // namespace cmgpu

namespace {

constexpr unsigned kScalars = 64;

int comm_setup(cmgpu_comm *m)
{
    CU(cudaSetDevice(m->device));
    CU(cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking));
    CU(cudaMalloc(&m->d_scalars, sizeof(double) * kScalars));
    CU(cudaMalloc(&m->d_counts, sizeof(unsigned) * (size_t)m->size));
    CU(cudaMallocHost(&m->h_counts, sizeof(unsigned) * (size_t)m->size));
    CU(cudaEventCreate(&m->ev0));
    CU(cudaEventCreate(&m->ev1));
    return CMGPU_OK;
}

int allreduce_f64(cmgpu_comm *m, double *values, unsigned n, ncclRedOp_t op)
{
    if (!m || !values)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if (n > kScalars)
        return fail(CMGPU_ERR_INVAL, "at most %u values", kScalars);
    if (!n)
        return CMGPU_OK;
    std::lock_guard<std::mutex> lk(m->mu);
    CU(cudaSetDevice(m->device));
    CU(cudaMemcpyAsync(m->d_scalars, values, sizeof(double) * n, cudaMemcpyHostToDevice, m->st));
    NC(ncclAllReduce(m->d_scalars, m->d_scalars, n, ncclDouble, op, m->nccl, m->st));
    CU(cudaMemcpyAsync(values, m->d_scalars, sizeof(double) * n, cudaMemcpyDeviceToHost, m->st));
    CU(cudaStreamSynchronize(m->st));
    return CMGPU_OK;
}

}  // namespace

extern "C" {

int cmgpu_comm_nccl_version(void)
{
    int v = 0;
    return ncclGetVersion(&v) == ncclSuccess ? v : 0;
}

int cmgpu_comm_unique_id(unsigned char id[CMGPU_COMM_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == CMGPU_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    if (!id)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    ncclUniqueId u;
    NC(ncclGetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return CMGPU_OK;
}

cmgpu_comm_t *cmgpu_comm_create(int device, int rank, int nranks, const unsigned char id[CMGPU_COMM_ID_BYTES])
{
    if (!id || nranks < 1 || rank < 0 || rank >= nranks) {
        fail(CMGPU_ERR_INVAL, "cmgpu_comm_create: rank %d of %d", rank, nranks);
        return nullptr;
    }
    cmgpu_comm *m = new (std::nothrow) cmgpu_comm;
    if (!m) {
        fail(CMGPU_ERR_NOMEM, "out of host memory");
        return nullptr;
    }
    m->device = device;
    m->rank = rank;
    m->size = nranks;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        fail(CMGPU_ERR_GENERIC, "cmgpu_comm_create: cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        delete m;
        return nullptr;
    }
    ncclResult_t r = ncclCommInitRank(&m->nccl, nranks, u, rank);
    if (r != ncclSuccess) {
        fail(CMGPU_ERR_GENERIC, "ncclCommInitRank(rank %d of %d): %s", rank, nranks, ncclGetErrorString(r));
        delete m;
        return nullptr;
    }
    if (comm_setup(m) != CMGPU_OK) {
        cmgpu_comm_destroy(m);
        return nullptr;
    }
    return m;
}

cmgpu_comm_t *cmgpu_comm_adopt(void *nccl_comm, int device)
{
    if (!nccl_comm) {
        fail(CMGPU_ERR_FAULT, "NULL communicator");
        return nullptr;
    }
    cmgpu_comm *m = new (std::nothrow) cmgpu_comm;
    if (!m) {
        fail(CMGPU_ERR_NOMEM, "out of host memory");
        return nullptr;
    }
    m->nccl = static_cast<ncclComm_t>(nccl_comm);
    m->owned = false;
    m->device = device;
    if (ncclCommUserRank(m->nccl, &m->rank) != ncclSuccess || ncclCommCount(m->nccl, &m->size) != ncclSuccess) {
        fail(CMGPU_ERR_GENERIC, "not a usable ncclComm_t");
        delete m;
        return nullptr;
    }
    if (comm_setup(m) != CMGPU_OK) {
        cmgpu_comm_destroy(m);
        return nullptr;
    }
    return m;
}

cmgpu_comm_t *cmgpu_comm_create_file(int device, int rank, int nranks, const char *path, int timeout_ms)
{
    if (!path || !*path) {
        fail(CMGPU_ERR_FAULT, "NULL path");
        return nullptr;
    }
    unsigned char id[CMGPU_COMM_ID_BYTES];
    if (rank == 0) {
        if (cmgpu_comm_unique_id(id) != CMGPU_OK)
            return nullptr;
        // publish atomically: a reader sees either no file or all 128 bytes
        const std::string tmp = std::string(path) + ".tmp";
        FILE *f = fopen(tmp.c_str(), "wb");
        if (!f || fwrite(id, 1, sizeof(id), f) != sizeof(id) || fclose(f) != 0 || rename(tmp.c_str(), path) != 0) {
            fail(CMGPU_ERR_GENERIC, "cannot publish the NCCL id in %s: %s", path, strerror(errno));
            return nullptr;
        }
    } else {
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms > 0 ? timeout_ms : 60000);
        for (;;) {
            FILE *f = fopen(path, "rb");
            if (f) {
                const size_t n = fread(id, 1, sizeof(id), f);
                fclose(f);
                if (n == sizeof(id))
                    break;
            }
            if (std::chrono::steady_clock::now() > deadline) {
                fail(CMGPU_ERR_GENERIC, "timed out waiting for the NCCL id in %s", path);
                return nullptr;
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(5));
        }
    }
    cmgpu_comm_t *m = cmgpu_comm_create(device, rank, nranks, id);
    // ncclCommInitRank returns once every rank has joined: the file has served its purpose
    if (rank == 0)
        unlink(path);
    return m;
}

void cmgpu_comm_destroy(cmgpu_comm_t *m)
{
    if (!m)
        return;
    cudaSetDevice(m->device);
    if (m->st)
        cudaStreamSynchronize(m->st);
    if (m->nccl && m->owned)
        ncclCommDestroy(m->nccl);
    if (m->st) cudaStreamDestroy(m->st);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    cudaFree(m->d_scalars);
    cudaFree(m->d_counts);
    cudaFree(m->d_gather);
    if (m->h_counts) cudaFreeHost(m->h_counts);
    if (m->h_gather) cudaFreeHost(m->h_gather);
    cudaGetLastError();
    delete m;
}

int cmgpu_comm_rank(const cmgpu_comm_t *m) { return m ? m->rank : -1; }
int cmgpu_comm_size(const cmgpu_comm_t *m) { return m ? m->size : 0; }
float cmgpu_comm_last_gather_ms(const cmgpu_comm_t *m) { return m ? m->last_ms : 0.f; }

int cmgpu_comm_barrier(cmgpu_comm_t *m)
{
    double one = 1.0;
    return allreduce_f64(m, &one, 1, ncclSum);
}

int cmgpu_comm_max(cmgpu_comm_t *m, double *values, unsigned n) { return allreduce_f64(m, values, n, ncclMax); }
int cmgpu_comm_sum(cmgpu_comm_t *m, double *values, unsigned n) { return allreduce_f64(m, values, n, ncclSum); }

int cmgpu_gather_results(cmgpu_ctx_t *c, cmgpu_comm_t *m, int root, uint32_t rate, int reset, cmgpu_result_t *results,
                         cmgpu_meter_state_t *states, int *rcs, unsigned *counts)
{
    CMGPU_TRACE("cmgpu_gather_results");
    if (!c || !m)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if (root < 0 || root >= m->size)
        return fail(CMGPU_ERR_INVAL, "root %d of %d ranks", root, m->size);
    if (c->device != m->device)
        return fail(CMGPU_ERR_INVAL, "context on device %d, communicator on device %d", c->device, m->device);
    const unsigned C = c->out_channels ? c->out_channels : c->channels;
    const unsigned row = c->row_u64;
    std::lock_guard<std::mutex> lkc(c->mu);
    std::lock_guard<std::mutex> lkm(m->mu);
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->cmp();
    CU(cudaEventRecord(m->ev0, st));

    // 1. my rows -> device staging (+ reset), ordered after every tick queued so far
    const unsigned mine = c->active;
    int rc = mine ? cmgpu::take_rows_locked(c, 0, mine, reset, false, rate) : cmgpu::ensure_take_buffers_locked(c);
    if (rc)
        return rc;

    // 2. how many streams each rank brings (the active count may differ and change between calls):
    //    every rank sends its count, then its rows; only the root waits on the host in between (it
    //    needs the counts to size its receives). The other ranks queue two sends and return.
    cmgpu::set_u32<<<1, 1, 0, st>>>(m->d_counts + m->rank, mine);
    CU(cudaGetLastError());
    if (m->rank == root) {
        NC(ncclGroupStart());
        ncclResult_t gr = ncclSuccess;
        for (int r = 0; r < m->size && gr == ncclSuccess; r++)
            if (r != root)
                gr = ncclRecv(m->d_counts + r, 1, ncclUint32, r, m->nccl, st);
        NC(ncclGroupEnd());                             // (closed before any error return: no group is left open)
        NC(gr);
        CU(cudaMemcpyAsync(m->h_counts, m->d_counts, sizeof(unsigned) * (size_t)m->size, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    } else {
        NC(ncclSend(m->d_counts + m->rank, 1, ncclUint32, root, m->nccl, st));
        if (mine)
            NC(ncclSend(c->d_take, (size_t)mine * row, ncclUint64, root, m->nccl, st));
        CU(cudaEventRecord(m->ev1, st));
        m->last_ms = 0.f;          // not waited for on this rank
        return CMGPU_OK;
    }
    size_t total = 0;
    for (int r = 0; r < m->size; r++)
        total += m->h_counts[r];

    // 3. rows to the root
    {
        const size_t need = total * row;
        if (need > m->gather_u64) {
            cudaFree(m->d_gather);
            if (m->h_gather)
                cudaFreeHost(m->h_gather);
            m->d_gather = nullptr;
            m->h_gather = nullptr;
            m->gather_u64 = 0;
            CU(cudaMalloc(&m->d_gather, need * sizeof(uint64_t)));
            CU(cudaMallocHost(&m->h_gather, need * sizeof(uint64_t)));
            m->gather_u64 = need;
        }
        size_t off = 0;
        NC(ncclGroupStart());
        ncclResult_t gr = ncclSuccess;
        for (int r = 0; r < m->size && gr == ncclSuccess; r++) {
            const size_t n = (size_t)m->h_counts[r] * row;
            if (r != root && n)
                gr = ncclRecv(m->d_gather + off, n, ncclUint64, r, m->nccl, st);
            off += n;
        }
        NC(ncclGroupEnd());
        NC(gr);
        off = 0;
        for (int r = 0; r < root; r++)
            off += (size_t)m->h_counts[r] * row;
        if (mine)
            CU(cudaMemcpyAsync(m->d_gather + off, c->d_take, sizeof(uint64_t) * mine * row, cudaMemcpyDeviceToDevice, st));
        if (need)
            CU(cudaMemcpyAsync(m->h_gather, m->d_gather, need * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaEventRecord(m->ev1, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1));

    // 4. decode + finalise on the root (host arithmetic: the reference's expression, the reference's libm)
    {
        if (counts)
            for (int r = 0; r < m->size; r++)
                counts[r] = m->h_counts[r];
        cmgpu::finalise_rows(m->h_gather, total, row, C, rate, results, states, rcs);
    }
    return CMGPU_OK;
}

}  // extern "C"
