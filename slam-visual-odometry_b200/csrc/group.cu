// Multi-GPU sequence path in C (include/b200vo.h, "device groups"): ONE sequence partitioned by contiguous
// frame-pair ranges over the GPUs of one node, one host thread per context, the 160-byte pose records
// all-gathered with NCCL over NVLink, chained by the caller on the host exactly as process_frame would
// (reference src/visual_odometry.cpp:340-374; the frame loop it replaces is src/main.cpp:128-194).
//
// NCCL is bound at first use with dlopen("libnccl.so.2"): a process that already carries an NCCL (torch's
// bundled one in the Python tests and bench) keeps exactly that copy, a plain C++ caller gets the system
// library.  Only five stable entry points are used.  There is no fallback: without NCCL a group of more
// than one device cannot be created.
//
// A device may be listed k times (adjacent entries, the same k for every device): it then carries k contexts, each on
// its own CUDA stream and host thread, with contiguous sub-ranges.  Every launch batch ends in latency-bound kernels
// (filter, deferred five-point, recoverPose) and every kernel in a partially filled last wave; a second stream fills
// those with the other context's throughput-bound kernels (bench.py MultiRunner: 28.0 -> 27.2 ms per 999 pairs with
// k = 2).  The contexts of a device write their records into ONE send buffer of the device, its first context (the
// leader) waits for the others through events and takes part in the NCCL all-gather, which runs over the distinct devices.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <thread>

#include "common.cuh"

namespace b200vo {
namespace {

typedef struct ncclComm* ncclComm_t;
struct NcclApi {
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int /* ncclDataType_t */, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    bool ok = false;
};
constexpr int kNcclUint8 = 1;       // ncclUint8 (nccl.h: ncclInt8 = 0, ncclUint8 = 1)

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return api;
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(h, "ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(h, "ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(h, "ncclGetVersion"));
    api.ok = api.CommInitAll && api.CommDestroy && api.AllGather && api.GroupStart && api.GroupEnd;
    return api;
}

}  // namespace
}  // namespace b200vo

using namespace b200vo;

struct b200vo_group {
    std::vector<b200vo_ctx*> ctx;                // n contexts, device-major: context i = k * (device index) + (index on the device)
    std::vector<int> devices;                    // [n] device of every context
    std::vector<int> udev;                       // [m] the distinct devices, m = n / k
    int k = 1;                                   // contexts per device
    std::vector<ncclComm_t> comm;                // [m] (empty when m == 1)
    std::vector<b200vo_pose_record*> d_send;     // [m] k * per_cap records: the ranges of the device's contexts (padded)
    std::vector<b200vo_pose_record*> d_recv;     // [m] n * per_cap gathered
    std::vector<cudaEvent_t> done;               // [n] recorded on a context's stream when its records are in d_send
    int per_cap = 0;
    int64_t launches_gather = 0;
};

static void contiguous_range(int total, int world, int rank, int* lo, int* hi) {     // first ranks get the remainder
    const int base = total / world, rem = total % world;
    *lo = rank * base + std::min(rank, rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
}

static int ensure_group_buffers(b200vo_group* g, int per) {
    if (per <= g->per_cap) return B200VO_OK;
    const int n = (int)g->ctx.size(), m = (int)g->udev.size();
    for (int i = 0; i < n; ++i) {
        B200VO_CUDA_CHECK(cudaSetDevice(g->devices[i]));
        B200VO_CUDA_CHECK(cudaStreamSynchronize(g->ctx[i]->stream));
    }
    for (int u = 0; u < m; ++u) {
        B200VO_CUDA_CHECK(cudaSetDevice(g->udev[u]));
        if (g->d_send[u]) cudaFree(g->d_send[u]);
        if (g->d_recv[u]) cudaFree(g->d_recv[u]);
        g->d_send[u] = g->d_recv[u] = nullptr;
        B200VO_CUDA_CHECK(cudaMalloc(&g->d_send[u], (size_t)per * g->k * sizeof(b200vo_pose_record)));
        B200VO_CUDA_CHECK(cudaMalloc(&g->d_recv[u], (size_t)per * n * sizeof(b200vo_pose_record)));
        B200VO_CUDA_CHECK(cudaMemset(g->d_send[u], 0, (size_t)per * g->k * sizeof(b200vo_pose_record)));
    }
    g->per_cap = per;
    return B200VO_OK;
}

extern "C" {

int b200vo_group_create(const b200vo_options* opt, const int* devices, int ndev, b200vo_group** out) {
    if (!opt || !devices || ndev < 1 || ndev > 64 || !out) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    // distinct devices in order of appearance; a device may be repeated k times in adjacent entries (k contexts on it)
    std::vector<int> udev;
    int k = 0;
    for (int i = 0; i < ndev; ) {
        int j = i;
        while (j < ndev && devices[j] == devices[i]) ++j;
        for (int u : udev) if (u == devices[i]) { set_error("device %d listed in two places (repeat a device in adjacent entries)", devices[i]); return B200VO_ERR_BAD_ARG; }
        if (k && j - i != k) { set_error("every device must be listed the same number of times"); return B200VO_ERR_BAD_ARG; }
        k = j - i;
        udev.push_back(devices[i]);
        i = j;
    }
    const int m = (int)udev.size();
    if (m > 1 && !nccl().ok) {
        set_error("NCCL (libnccl.so.2) could not be loaded: a group of %d devices needs it (no host-gather fallback)", m);
        return B200VO_ERR_CUDA;
    }
    b200vo_group* g = new b200vo_group();
    g->devices.assign(devices, devices + ndev);
    g->udev = udev;
    g->k = k;
    g->d_send.assign(m, nullptr);
    g->d_recv.assign(m, nullptr);
    g->done.assign(ndev, nullptr);
    for (int i = 0; i < ndev; ++i) {
        b200vo_ctx* c = nullptr;
        const int rc = b200vo_create(opt, devices[i], &c);
        if (rc) { b200vo_group_destroy(g); return rc; }
        g->ctx.push_back(c);
        if (cudaEventCreateWithFlags(&g->done[i], cudaEventDisableTiming) != cudaSuccess) { set_error("cudaEventCreate failed"); b200vo_group_destroy(g); return B200VO_ERR_CUDA; }
    }
    if (m > 1) {
        g->comm.assign(m, nullptr);
        const int rc = nccl().CommInitAll(g->comm.data(), m, udev.data());
        if (rc != 0) {
            set_error("ncclCommInitAll failed: %s", nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
            g->comm.clear();
            b200vo_group_destroy(g);
            return B200VO_ERR_CUDA;
        }
    }
    *out = g;
    return B200VO_OK;
}

void b200vo_group_destroy(b200vo_group* g) {
    if (!g) return;
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        cudaSetDevice(g->devices[i]);
        if (g->ctx[i] && g->ctx[i]->stream) cudaStreamSynchronize(g->ctx[i]->stream);
    }
    for (size_t u = 0; u < g->udev.size(); ++u) {
        cudaSetDevice(g->udev[u]);
        if (u < g->comm.size() && g->comm[u]) nccl().CommDestroy(g->comm[u]);
        if (g->d_send[u]) cudaFree(g->d_send[u]);
        if (g->d_recv[u]) cudaFree(g->d_recv[u]);
    }
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        cudaSetDevice(g->devices[i]);
        if (i < g->done.size() && g->done[i]) cudaEventDestroy(g->done[i]);
        b200vo_destroy(g->ctx[i]);
    }
    delete g;
}

int b200vo_group_size(const b200vo_group* g) { return g ? (int)g->ctx.size() : B200VO_ERR_BAD_ARG; }
b200vo_ctx* b200vo_group_ctx(b200vo_group* g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }

int b200vo_partition_pairs(int total_pairs, int world, int rank, int* lo, int* hi) {
    if (total_pairs < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    contiguous_range(total_pairs, world, rank, lo, hi);
    return B200VO_OK;
}

int b200vo_nccl_version(void) {
    int v = 0;
    if (nccl().ok && nccl().GetVersion) nccl().GetVersion(&v);
    return v;
}

// All-gather of the per-device record ranges: device i contributes counts[i] records from its d_send buffer
// (filled by b200vo_group_process_sequence, or by the caller through b200vo_group_send_buffer); every rank sends
// the same padded count, as ncclAllGather requires.  One NCCL group call issued from this thread on the contexts'
// own streams; the gathered records of device 0 are copied to `out` in global pair order.
int b200vo_group_allgather_records(b200vo_group* g, const int* counts, b200vo_pose_record* out) {
    if (!g || !counts || !out) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    const int n = (int)g->ctx.size(), m = (int)g->udev.size(), k = g->k;
    int per = 0, total = 0;
    for (int i = 0; i < n; ++i) { if (counts[i] < 0) { set_error("negative count"); return B200VO_ERR_BAD_ARG; } per = std::max(per, counts[i]); total += counts[i]; }
    if (per > g->per_cap) { set_error("counts exceed the group's record buffers (%d > %d)", per, g->per_cap); return B200VO_ERR_CAPACITY; }
    if (total == 0) return B200VO_OK;
    per = g->per_cap;                     // the layout of d_send / d_recv: context i's records start at i * per_cap
    // join: the leader stream of every device waits until the device's other contexts have written their records
    for (int i = 0; i < n; ++i) {
        if (i % k == 0) continue;
        B200VO_CUDA_CHECK(cudaSetDevice(g->devices[i]));
        B200VO_CUDA_CHECK(cudaEventRecord(g->done[i], g->ctx[i]->stream));
        B200VO_CUDA_CHECK(cudaStreamWaitEvent(g->ctx[i - i % k]->stream, g->done[i], 0));
    }
    const size_t bytes = (size_t)per * k * sizeof(b200vo_pose_record);
    if (m > 1) {
        NcclApi& nc = nccl();
        int rc = nc.GroupStart();
        for (int u = 0; u < m && rc == 0; ++u)
            rc = nc.AllGather(g->d_send[u], g->d_recv[u], bytes, kNcclUint8, g->comm[u], g->ctx[(size_t)u * k]->stream);
        const int rc2 = nc.GroupEnd();
        if (rc == 0) rc = rc2;
        if (rc != 0) { set_error("ncclAllGather failed: %s", nc.GetErrorString ? nc.GetErrorString(rc) : "?"); return B200VO_ERR_CUDA; }
        g->launches_gather += m;
    }
    // device 0's copy -> host, dropping the padding
    B200VO_CUDA_CHECK(cudaSetDevice(g->devices[0]));
    const b200vo_pose_record* src = (m > 1) ? g->d_recv[0] : g->d_send[0];
    int o = 0;
    for (int i = 0; i < n; ++i) {
        if (counts[i])
            B200VO_CUDA_CHECK(cudaMemcpyAsync(out + o, src + (size_t)i * per, (size_t)counts[i] * sizeof(b200vo_pose_record), cudaMemcpyDeviceToHost, g->ctx[0]->stream));
        o += counts[i];
    }
    B200VO_CUDA_CHECK(cudaStreamSynchronize(g->ctx[0]->stream));
    for (int i = 1; i < n; ++i) {         // the other ranks' collectives have completed when rank 0's has; drain their streams anyway
        B200VO_CUDA_CHECK(cudaSetDevice(g->devices[i]));
        B200VO_CUDA_CHECK(cudaStreamSynchronize(g->ctx[i]->stream));
    }
    return B200VO_OK;
}

// The whole multi-GPU sequence step: pairs (i, i+1), i < nframes-1, with pair ids first_pair_id + i.  Device r owns
// the contiguous range b200vo_partition_pairs gives it and therefore frames lo..hi (a boundary frame is detected
// on both neighbours); one host thread per device runs upload -> detect -> process_pairs in launch batches of
// batch_frames on its own context (contexts are not thread-safe; each thread touches only its own).
int b200vo_group_process_sequence(b200vo_group* g, const uint8_t* frames, int nframes, size_t pitch, size_t frame_stride,
                                  int channels, int batch_frames, int64_t first_pair_id, b200vo_pose_record* out) {
    if (!g || !frames || nframes < 2 || !out || batch_frames < 1) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    const int n = (int)g->ctx.size();
    const int P = nframes - 1;
    for (int i = 0; i < n; ++i)
        if (batch_frames + 1 > g->ctx[i]->store.slots) { set_error("batch_frames %d needs max_frames >= %d", batch_frames, batch_frames + 1); return B200VO_ERR_BAD_ARG; }
    int rc = ensure_group_buffers(g, (P + n - 1) / n);
    if (rc) return rc;
    std::vector<int> counts((size_t)n), status((size_t)n, B200VO_OK);
    std::vector<std::string> errors((size_t)n);
    auto worker = [&](int r) {
        int lo, hi;
        contiguous_range(P, n, r, &lo, &hi);
        counts[(size_t)r] = hi - lo;
        if (hi == lo) return;
        b200vo_ctx* c = g->ctx[(size_t)r];
        const int slots = c->store.slots, B = std::min(batch_frames, slots - 1);
        auto fail = [&](int code) { status[(size_t)r] = code; errors[(size_t)r] = b200vo_last_error(); };
        if (cudaSetDevice(g->devices[(size_t)r]) != cudaSuccess) { set_error("cudaSetDevice failed"); fail(B200VO_ERR_CUDA); return; }
        const uint8_t* base = frames + (size_t)lo * frame_stride;       // this rank's frames lo .. hi
        const int F = hi - lo + 1;
        int done_f = 0, done_p = 0, prev = 0, e;
        while (done_f < F) {
            const int nb = std::min(B, F - done_f);
            const int first = done_f == 0 ? 0 : (prev + 1) % slots;
            int up = 0;
            while (up < nb) {                                            // the ring may wrap inside a batch
                const int s = (first + up) % slots, run = std::min(nb - up, slots - s);
                if ((e = b200vo_upload_frames(c, base + (size_t)(done_f + up) * frame_stride, run, pitch, frame_stride, channels, s))) { fail(e); return; }
                up += run;
            }
            if ((e = b200vo_detect(c, first, nb))) { fail(e); return; }
            const int p0 = done_f == 0 ? first : prev, np = done_f == 0 ? nb - 1 : nb;
            if (np > 0) {
                void* dev = nullptr;
                if ((e = b200vo_process_pairs_device(c, p0, np, first_pair_id + lo + done_p, &dev))) { fail(e); return; }
                if (cudaMemcpyAsync(g->d_send[(size_t)(r / g->k)] + (size_t)(r % g->k) * g->per_cap + done_p, dev, (size_t)np * sizeof(b200vo_pose_record), cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) {
                    set_error("record copy failed"); fail(B200VO_ERR_CUDA); return;
                }
            }
            prev = (first + nb - 1) % slots;
            done_f += nb;
            done_p += np;
        }
    };
    if (n == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int r = 0; r < n; ++r) th.emplace_back(worker, r);
        for (auto& t : th) t.join();
    }
    for (int r = 0; r < n; ++r)
        if (status[(size_t)r]) { set_error("device %d: %s", g->devices[(size_t)r], errors[(size_t)r].c_str()); return status[(size_t)r]; }
    return b200vo_group_allgather_records(g, counts.data(), out);
}

}  // extern "C"
