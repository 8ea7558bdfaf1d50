// fake_nccl.cpp -- in-process stand-in for the handful of NCCL calls sepgpu_dd.cu makes.  TEST INFRASTRUCTURE ONLY,
// part of tests/_build/libsep_emu.so (the library finds it because the emulated build opens itself where the
// product opens libnccl.so.2).
//
// "Ranks" are host THREADS of one process, each driving its own sepgpu_ctx on the CPU kernel emulator.  Point-to-point
// sends are buffered (copied into a FIFO per (source, destination) pair, matched with receives in issue order, which is
// NCCL's matching rule); receives block until the message is there.  Collectives meet at a generation barrier and
// reduce in rank order.  Streams are synchronous in the emulator, so every call completes before it returns.
#include "nccl.h"

#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {
struct World {
    int nranks = 0, joined = 0;
    std::mutex m;
    std::condition_variable cv;
    std::map<std::pair<int, int>, std::deque<std::vector<unsigned char>>> box;     // (src, dst) -> messages
    // collectives
    int arrived = 0;
    long long gen = 0;
    std::vector<std::vector<unsigned char>> slot;
};
struct Comm { World *w; int rank; };
std::mutex g_m;
std::map<std::string, World *> g_worlds;
long long g_ids = 0;

size_t type_size(ncclDataType_t t)
{
    switch (t) {
    case ncclInt8: case ncclUint8: return 1;
    case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
    case ncclInt64: case ncclUint64: case ncclFloat64: return 8;
    default: return 0;
    }
}

// every rank deposits `bytes` bytes, all wait, `fn(slots)` runs on every rank, all wait again
template <class F> void collective(Comm *c, const void *send, size_t bytes, F fn)
{
    World *w = c->w;
    std::unique_lock<std::mutex> lk(w->m);
    w->slot[c->rank].assign((const unsigned char *)send, (const unsigned char *)send + bytes);
    long long g = w->gen;
    if (++w->arrived == w->nranks) { w->arrived = 0; w->gen++; w->cv.notify_all(); }
    else w->cv.wait(lk, [&] { return w->gen != g; });
    fn(w->slot);
    g = w->gen;
    if (++w->arrived == w->nranks) { w->arrived = 0; w->gen++; w->cv.notify_all(); }
    else w->cv.wait(lk, [&] { return w->gen != g; });
}
}  // namespace

extern "C" {
ncclResult_t ncclGetUniqueId(ncclUniqueId *id)
{
    std::lock_guard<std::mutex> g(g_m);
    memset(id, 0, sizeof *id);
    snprintf(id->internal, sizeof id->internal, "sepgpu-emu-world-%lld", ++g_ids);
    return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank)
{
    World *w;
    {
        std::lock_guard<std::mutex> g(g_m);
        std::string key(id.internal, sizeof id.internal);
        auto it = g_worlds.find(key);
        if (it == g_worlds.end()) { w = new World(); w->nranks = nranks; w->slot.resize(nranks); g_worlds[key] = w; }
        else w = it->second;
    }
    if (w->nranks != nranks || rank < 0 || rank >= nranks) return ncclInternalError;
    std::unique_lock<std::mutex> lk(w->m);
    w->joined++;
    w->cv.notify_all();
    w->cv.wait(lk, [&] { return w->joined >= w->nranks; });
    Comm *c = new Comm{w, rank};
    *comm = (ncclComm_t)c;
    return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm) { delete (Comm *)comm; return ncclSuccess; }
ncclResult_t ncclGroupStart(void) { return ncclSuccess; }
ncclResult_t ncclGroupEnd(void) { return ncclSuccess; }
const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "fake NCCL: internal error"; }

ncclResult_t ncclSend(const void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t comm, cudaStream_t)
{
    Comm *c = (Comm *)comm;
    const size_t bytes = count * type_size(t);
    std::lock_guard<std::mutex> lk(c->w->m);
    c->w->box[{c->rank, peer}].emplace_back((const unsigned char *)buf, (const unsigned char *)buf + bytes);
    c->w->cv.notify_all();
    return ncclSuccess;
}

ncclResult_t ncclRecv(void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t comm, cudaStream_t)
{
    Comm *c = (Comm *)comm;
    const size_t bytes = count * type_size(t);
    std::unique_lock<std::mutex> lk(c->w->m);
    auto &q = c->w->box[{peer, c->rank}];
    c->w->cv.wait(lk, [&] { return !q.empty(); });
    if (q.front().size() != bytes) return ncclInternalError;         // a size mismatch would hang real NCCL
    memcpy(buf, q.front().data(), bytes);
    q.pop_front();
    return ncclSuccess;
}

ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t comm, cudaStream_t)
{
    if (t != ncclFloat64 || (op != ncclSum && op != ncclMax)) return ncclInternalError;
    Comm *c = (Comm *)comm;
    std::vector<double> out(count);
    collective(c, send, count * sizeof(double), [&](std::vector<std::vector<unsigned char>> &slots) {
        for (size_t k = 0; k < count; k++) {
            double acc = ((const double *)slots[0].data())[k];
            for (int r = 1; r < c->w->nranks; r++) {
                const double v = ((const double *)slots[r].data())[k];
                acc = op == ncclSum ? acc + v : (v > acc ? v : acc);
            }
            out[k] = acc;
        }
    });
    memcpy(recv, out.data(), count * sizeof(double));
    return ncclSuccess;
}

ncclResult_t ncclAllGather(const void *send, void *recv, size_t sendcount, ncclDataType_t t, ncclComm_t comm, cudaStream_t)
{
    Comm *c = (Comm *)comm;
    const size_t bytes = sendcount * type_size(t);
    std::vector<unsigned char> out(bytes * c->w->nranks);
    collective(c, send, bytes, [&](std::vector<std::vector<unsigned char>> &slots) {
        for (int r = 0; r < c->w->nranks; r++) memcpy(out.data() + bytes * r, slots[r].data(), bytes);
    });
    memcpy(recv, out.data(), out.size());
    return ncclSuccess;
}
}
