// tests/support/hostemu/fake_nccl.cpp — TEST INFRASTRUCTURE ONLY.
//
// A stand-in for libnccl.so.2 (the five entry points csrc/lh_soil_api.cu resolves with dlopen / dlsym) for the CPU-only run of
// the multi-rank path on the emulated build: the "ranks" are host threads of one process, each driving its own ctx over its own
// column shard; ncclCommInitRank is a rendezvous of the threads that hold the same unique id, ncclAllReduce sums their buffers
// in rank order.  The test process loads this library (soname libnccl.so.2) BEFORE the product library first asks for NCCL, so
// the product's own `dlopen("libnccl.so.2", RTLD_NOLOAD)` finds it — the product code is unchanged and unaware.
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {
struct Group {
    int nranks = 0, joined = 0;
    std::mutex m;
    std::condition_variable cv;
    // all-reduce rendezvous
    std::vector<const void*> send;
    std::vector<void*> recv;
    int arrived = 0, left = 0;
    uint64_t round = 0;
};
struct Comm { Group* g; int rank; };
std::mutex g_m;
std::map<std::string, Group*> g_groups;
uint64_t g_next_id = 1;
}  // namespace

extern "C" {
struct ncclUniqueId { char internal[128]; };

int ncclGetUniqueId(ncclUniqueId* id)
{
    std::lock_guard<std::mutex> lock(g_m);
    memset(id->internal, 0, sizeof id->internal);
    const uint64_t v = g_next_id++;
    memcpy(id->internal, "FAKENCCL", 8);
    memcpy(id->internal + 8, &v, sizeof v);
    return 0;
}

int ncclCommInitRank(Comm** comm, int nranks, ncclUniqueId id, int rank)
{
    if (!comm || nranks < 1 || rank < 0 || rank >= nranks || memcmp(id.internal, "FAKENCCL", 8) != 0) return 4;   // ncclInvalidArgument
    Group* g;
    {
        std::lock_guard<std::mutex> lock(g_m);
        Group*& slot = g_groups[std::string(id.internal, 128)];
        if (!slot) { slot = new Group(); slot->nranks = nranks; slot->send.resize(nranks); slot->recv.resize(nranks); }
        g = slot;
    }
    if (g->nranks != nranks) return 4;
    std::unique_lock<std::mutex> lock(g->m);
    ++g->joined;
    g->cv.notify_all();
    g->cv.wait(lock, [&] { return g->joined >= g->nranks; });     // every rank of the communicator must call (as with NCCL)
    *comm = new Comm{g, rank};
    return 0;
}

// count doubles (datatype 8 = ncclDouble), op 0 = ncclSum; the stream argument is ignored: the emulated build is synchronous here
int ncclAllReduce(const void* send, void* recv, size_t count, int datatype, int op, Comm* comm, void* /*stream*/)
{
    if (!comm || datatype != 8 || op != 0) return 4;
    Group* g = comm->g;
    std::unique_lock<std::mutex> lock(g->m);
    const uint64_t round = g->round;
    g->send[comm->rank] = send;
    g->recv[comm->rank] = recv;
    if (++g->arrived == g->nranks) {
        std::vector<double> sum(count, 0.0);
        for (int r = 0; r < g->nranks; ++r)                       // rank order: the same bits on every rank
            for (size_t i = 0; i < count; ++i) sum[i] += ((const double*)g->send[r])[i];
        for (int r = 0; r < g->nranks; ++r) memcpy(g->recv[r], sum.data(), count * sizeof(double));
        g->arrived = 0;
        ++g->round;
        g->cv.notify_all();
    } else {
        g->cv.wait(lock, [&] { return g->round != round; });
    }
    return 0;
}

int ncclCommDestroy(Comm* comm)
{
    delete comm;
    return 0;
}

const char* ncclGetErrorString(int e) { return e == 0 ? "no error" : e == 4 ? "invalid argument" : "fake NCCL error"; }
}
