// dcr_comm.cu — life cycle of a dcr_comm (include/dcr.h): allocation of the IPC-exportable result buffer, handle exchange,
// mapping of the peers' buffers.  The device side lives in dcr_comm.cuh.
#include <algorithm>

#include "dcr_comm.cuh"

using namespace dcr;

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int dcr_comm_create(int rank, int world, int64_t n_edges, dcr_comm** out) {
    if (!out || world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world || n_edges < 0) {
        set_error("dcr_comm_create: bad arguments (world <= %d)", COMM_MAX_WORLD);
        return 1;
    }
    dcr_comm* c = new dcr_comm();
    c->rank = rank; c->world = world; c->device = current_device();
    c->n_edges = n_edges;
    c->chunk = std::max<int64_t>(4, (n_edges + 3) / 4 * 4);
    c->flag_off = align_up((size_t)c->chunk * 24, 256);
    c->bytes = c->flag_off + align_up(sizeof(CommFlags), 256);
    c->epoch = 0;
    c->connected = (world == 1);
    c->local = nullptr; c->d_peers = nullptr;
    for (int p = 0; p < COMM_MAX_WORLD; ++p) c->peer[p] = nullptr;
    cudaError_t e = cudaMalloc((void**)&c->local, c->bytes);            // plain cudaMalloc: exportable through CUDA IPC
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, c->bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_peers, sizeof(unsigned char*) * COMM_MAX_WORLD);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "dcr_comm_create", __FILE__, __LINE__); }
    c->peer[rank] = c->local;
    e = cudaMemcpy(c->d_peers, c->peer, sizeof(unsigned char*) * COMM_MAX_WORLD, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "dcr_comm_create", __FILE__, __LINE__); }
    *out = c;
    return 0;
}

extern "C" int dcr_comm_handle(dcr_comm* c, void* handle64_host) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!c || !handle64_host) { set_error("dcr_comm_handle: NULL argument"); return 1; }
    DCR_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64_host, c->local));
    return 0;
}

extern "C" int dcr_comm_connect(dcr_comm* c, const void* handles_host) {
    if (!c || !handles_host) { set_error("dcr_comm_connect: NULL argument"); return 1; }
    const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)handles_host;
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) continue;
        void* ptr = nullptr;
        DCR_CUDA(cudaIpcOpenMemHandle(&ptr, h[p], cudaIpcMemLazyEnablePeerAccess));
        c->peer[p] = (unsigned char*)ptr;
    }
    DCR_CUDA(cudaMemcpy(c->d_peers, c->peer, sizeof(unsigned char*) * COMM_MAX_WORLD, cudaMemcpyHostToDevice));
    c->connected = true;
    return 0;
}

extern "C" void* dcr_comm_buffer(dcr_comm* c) { return c ? c->local : nullptr; }
extern "C" int64_t dcr_comm_chunk(dcr_comm* c) { return c ? c->chunk : 0; }

extern "C" int dcr_comm_error(dcr_comm* c) {                               // synchronises
    if (!c) return 1;
    unsigned int err = 0;
    if (cudaMemcpy(&err, c->local + c->flag_off + offsetof(CommFlags, error), sizeof(err), cudaMemcpyDeviceToHost) != cudaSuccess)
        return 1;
    return (int)err;
}

extern "C" int dcr_comm_destroy(dcr_comm* c) {
    if (!c) return 0;
    cudaDeviceSynchronize();
    for (int p = 0; p < c->world; ++p)
        if (p != c->rank && c->peer[p]) cudaIpcCloseMemHandle(c->peer[p]);
    if (c->d_peers) cudaFree(c->d_peers);
    if (c->local) cudaFree(c->local);
    delete c;
    return 0;
}

