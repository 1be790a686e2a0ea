/*
 * oracle/mpi_stub/mpi_threads.cpp -- TEST INFRASTRUCTURE ONLY.
 * Thread-backed implementation of the MPI subset declared in mpi.h: every rank is a std::thread,
 * point-to-point messages are eager copies through a mailbox, collectives are built on top of them.
 * One global mutex/condvar -- performance is irrelevant, determinism and simplicity are the point.
 * Reductions are summed in ascending comm-rank order.
 */
#include "mpi.h"

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

namespace {
std::mutex g_mu;
std::condition_variable g_cv;

struct CommShared {
    std::vector<int> members;  // world ranks, comm-rank order
    std::vector<int> dims;     // cartesian dims (empty: not a cart comm)
    std::map<std::tuple<int, int, int>, std::deque<std::vector<char>>> mail;  // (src,dst,tag)
    int bar_count = 0;
    long bar_gen = 0;
    std::map<std::pair<int, long>, std::shared_ptr<CommShared>> children;  // (seq, color)
    std::map<int, std::vector<void*>> winbases;
};
}  // namespace

struct stub_comm {
    std::shared_ptr<CommShared> sh;
    int rank = 0;
    int child_seq = 0;
    int win_seq = 0;
};
struct stub_win {
    std::shared_ptr<CommShared> sh;
    stub_comm* comm;
    int seq;
    int disp_unit;
};
struct SubRecv {
    void* buf;
    size_t cap;
    int src, tag;
    bool done;
    size_t got;
};
struct stub_req {
    stub_comm* comm = nullptr;
    std::vector<SubRecv> subs;  // empty => already complete (eager send)
};

namespace {
thread_local stub_comm* t_world = nullptr;

inline int dt_size(MPI_Datatype d) { return d & 0xff; }
enum { TAG_RED = -101, TAG_BCAST = -102, TAG_SCAT = -103, TAG_GATH = -104, TAG_ALLG = -105 };

void post_send(stub_comm* c, int dst, int tag, const void* buf, size_t bytes) {
    std::vector<char> m(bytes);
    if (bytes) std::memcpy(m.data(), buf, bytes);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        c->sh->mail[std::make_tuple(c->rank, dst, tag)].push_back(std::move(m));
    }
    g_cv.notify_all();
}
// caller holds g_mu
bool try_recv_locked(stub_comm* c, SubRecv& s) {
    auto it = c->sh->mail.find(std::make_tuple(s.src, c->rank, s.tag));
    if (it == c->sh->mail.end() || it->second.empty()) return false;
    auto& m = it->second.front();
    s.got = m.size();
    if (m.size() > s.cap) {
        std::fprintf(stderr, "[mpi_stub] message truncated: %zu > %zu\n", m.size(), s.cap);
        std::abort();
    }
    if (!m.empty()) std::memcpy(s.buf, m.data(), m.size());
    it->second.pop_front();
    s.done = true;
    return true;
}
void blocking_recv(stub_comm* c, int src, int tag, void* buf, size_t cap, size_t* got = nullptr) {
    SubRecv s{buf, cap, src, tag, false, 0};
    std::unique_lock<std::mutex> lk(g_mu);
    g_cv.wait(lk, [&] { return try_recv_locked(c, s); });
    if (got) *got = s.got;
}
void barrier(stub_comm* c) {
    std::unique_lock<std::mutex> lk(g_mu);
    auto& sh = *c->sh;
    long gen = sh.bar_gen;
    if (++sh.bar_count == (int)sh.members.size()) {
        sh.bar_count = 0;
        sh.bar_gen++;
        g_cv.notify_all();
    } else {
        g_cv.wait(lk, [&] { return sh.bar_gen != gen; });
    }
}
std::vector<int> coords_of(const std::vector<int>& dims, int rank) {
    std::vector<int> c(dims.size());
    for (int d = (int)dims.size() - 1; d >= 0; --d) {
        c[d] = rank % dims[d];
        rank /= dims[d];
    }
    return c;
}
int rank_of(const std::vector<int>& dims, const int* coords) {
    int r = 0;
    for (size_t d = 0; d < dims.size(); ++d) r = r * dims[d] + coords[d];
    return r;
}
// collective child creation: all members call with identical (members, dims) for their colour
stub_comm* make_child(stub_comm* parent, long color, const std::vector<int>& members_parent_ranks,
                      const std::vector<int>& dims) {
    int seq = parent->child_seq++;
    std::shared_ptr<CommShared> sh;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto key = std::make_pair(seq, color);
        auto it = parent->sh->children.find(key);
        if (it == parent->sh->children.end()) {
            sh = std::make_shared<CommShared>();
            for (int pr : members_parent_ranks) sh->members.push_back(parent->sh->members[pr]);
            sh->dims = dims;
            parent->sh->children[key] = sh;
        } else {
            sh = it->second;
        }
    }
    auto* c = new stub_comm;
    c->sh = sh;
    int me = parent->rank;
    c->rank = (int)(std::find(members_parent_ranks.begin(), members_parent_ranks.end(), me) -
                    members_parent_ranks.begin());
    return c;
}
}  // namespace

extern "C" {

MPI_Comm stub_comm_world(void) {
    if (!t_world) {  // single-rank use without stub_mpi_run
        auto* c = new stub_comm;
        c->sh = std::make_shared<CommShared>();
        c->sh->members = {0};
        t_world = c;
    }
    return t_world;
}

void stub_mpi_run(int nranks, void (*fn)(int, void*), void* arg) {
    auto sh = std::make_shared<CommShared>();
    for (int r = 0; r < nranks; ++r) sh->members.push_back(r);
    std::vector<std::thread> th;
    for (int r = 0; r < nranks; ++r) {
        th.emplace_back([=] {
            auto* c = new stub_comm;
            c->sh = sh;
            c->rank = r;
            t_world = c;
            fn(r, arg);
            t_world = nullptr;
            delete c;
        });
    }
    for (auto& t : th) t.join();
}

int MPI_Init(int*, char***) { return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm, int code) { std::exit(code ? code : 1); }
int MPI_Comm_rank(MPI_Comm c, int* r) { *r = c->rank; return 0; }
int MPI_Comm_size(MPI_Comm c, int* s) { *s = (int)c->sh->members.size(); return 0; }
int MPI_Comm_free(MPI_Comm* c) {
    if (*c && *c != t_world) delete *c;
    *c = MPI_COMM_NULL;
    return 0;
}
int MPI_Barrier(MPI_Comm c) { barrier(c); return 0; }

int MPI_Cart_create(MPI_Comm c, int nd, const int* dims, const int*, int, MPI_Comm* out) {
    int n = 1;
    std::vector<int> d(dims, dims + nd);
    for (int x : d) n *= x;
    int sz = (int)c->sh->members.size();
    if (n > sz) { std::fprintf(stderr, "[mpi_stub] cart larger than comm\n"); std::abort(); }
    std::vector<int> mem(n);
    for (int i = 0; i < n; ++i) mem[i] = i;
    if (c->rank >= n) {  // not part of the grid
        c->child_seq++;
        *out = MPI_COMM_NULL;
        return 0;
    }
    *out = make_child(c, 0, mem, d);
    return 0;
}
int MPI_Cart_sub(MPI_Comm c, const int* remain, MPI_Comm* out) {
    const auto& dims = c->sh->dims;
    int nd = (int)dims.size();
    auto mine = coords_of(dims, c->rank);
    std::vector<int> sub_dims, mem;
    long color = 0;
    for (int d = 0; d < nd; ++d) {
        if (remain[d]) sub_dims.push_back(dims[d]);
        else color = color * (dims[d] + 1) + mine[d] + 1;
    }
    int n = (int)c->sh->members.size();
    for (int r = 0; r < n; ++r) {
        auto cr = coords_of(dims, r);
        bool same = true;
        for (int d = 0; d < nd; ++d)
            if (!remain[d] && cr[d] != mine[d]) same = false;
        if (same) mem.push_back(r);
    }
    *out = make_child(c, color, mem, sub_dims);
    return 0;
}
int MPI_Cart_coords(MPI_Comm c, int rank, int maxd, int* coords) {
    auto cr = coords_of(c->sh->dims, rank);
    for (int d = 0; d < maxd && d < (int)cr.size(); ++d) coords[d] = cr[d];
    return 0;
}
int MPI_Cart_rank(MPI_Comm c, const int* coords, int* rank) { *rank = rank_of(c->sh->dims, coords); return 0; }
int MPI_Cart_get(MPI_Comm c, int maxd, int* dims, int* periods, int* coords) {
    auto cr = coords_of(c->sh->dims, c->rank);
    for (int d = 0; d < maxd && d < (int)cr.size(); ++d) {
        dims[d] = c->sh->dims[d];
        periods[d] = 0;
        coords[d] = cr[d];
    }
    return 0;
}
int MPI_Comm_group(MPI_Comm, MPI_Group* g) { *g = 1; return 0; }
int MPI_Group_incl(MPI_Group, int, const int*, MPI_Group* g) { *g = 1; return 0; }
int MPI_Group_free(MPI_Group* g) { *g = 0; return 0; }
int MPI_Comm_create_group(MPI_Comm, MPI_Group, int, MPI_Comm*) {
    std::fprintf(stderr, "[mpi_stub] MPI_Comm_create_group is not on the LU path\n");
    std::abort();
}

int MPI_Isend(const void* buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm c, MPI_Request* rq) {
    post_send(c, dest, tag, buf, (size_t)count * dt_size(dt));
    *rq = new stub_req;
    (*rq)->comm = c;
    return 0;
}
int MPI_Irecv(void* buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Request* rq) {
    auto* r = new stub_req;
    r->comm = c;
    r->subs.push_back(SubRecv{buf, (size_t)count * dt_size(dt), src, tag, false, 0});
    *rq = r;
    return 0;
}
static void finish(stub_req* r, MPI_Status* st) {
    if (st && !r->subs.empty()) {
        st->MPI_SOURCE = r->subs[0].src;
        st->MPI_TAG = r->subs[0].tag;
        st->MPI_ERROR = 0;
        st->count_bytes = (int)r->subs[0].got;
    }
}
int MPI_Wait(MPI_Request* rq, MPI_Status* st) {
    if (!rq || !*rq) return 0;
    stub_req* r = *rq;
    {
        std::unique_lock<std::mutex> lk(g_mu);
        for (auto& s : r->subs)
            if (!s.done) g_cv.wait(lk, [&] { return try_recv_locked(r->comm, s); });
    }
    finish(r, st);
    delete r;
    *rq = MPI_REQUEST_NULL;
    return 0;
}
int MPI_Waitall(int n, MPI_Request* rq, MPI_Status* st) {
    for (int i = 0; i < n; ++i) MPI_Wait(&rq[i], st ? &st[i] : nullptr);
    return 0;
}
int MPI_Waitany(int n, MPI_Request* rq, int* idx, MPI_Status* st) {
    std::unique_lock<std::mutex> lk(g_mu);
    int found = -1;
    bool any = false;
    for (int i = 0; i < n; ++i) any = any || rq[i];
    if (!any) { *idx = MPI_UNDEFINED; return 0; }
    g_cv.wait(lk, [&] {
        for (int i = 0; i < n; ++i) {
            if (!rq[i]) continue;
            bool all = true;
            for (auto& s : rq[i]->subs)
                if (!s.done && !try_recv_locked(rq[i]->comm, s)) all = false;
            if (all) { found = i; return true; }
        }
        return false;
    });
    lk.unlock();
    finish(rq[found], st);
    delete rq[found];
    rq[found] = MPI_REQUEST_NULL;
    *idx = found;
    return 0;
}
int MPI_Request_free(MPI_Request* rq) {
    if (rq && *rq) { delete *rq; *rq = MPI_REQUEST_NULL; }
    return 0;
}
int MPI_Get_count(const MPI_Status* st, MPI_Datatype dt, int* count) { *count = st->count_bytes / dt_size(dt); return 0; }

int MPI_Sendrecv(const void* sbuf, int scount, MPI_Datatype sdt, int dest, int stag, void* rbuf, int rcount,
                 MPI_Datatype rdt, int src, int rtag, MPI_Comm c, MPI_Status* st) {
    // eager send first (copies sbuf before rbuf may alias/overwrite it)
    post_send(c, dest, stag, sbuf, (size_t)scount * dt_size(sdt));
    size_t got = 0;
    blocking_recv(c, src, rtag, rbuf, (size_t)rcount * dt_size(rdt), &got);
    if (st) { st->MPI_SOURCE = src; st->MPI_TAG = rtag; st->MPI_ERROR = 0; st->count_bytes = (int)got; }
    return 0;
}

int MPI_Reduce(const void* sbuf, void* rbuf, int count, MPI_Datatype dt, MPI_Op, int root, MPI_Comm c) {
    size_t bytes = (size_t)count * dt_size(dt);
    int n = (int)c->sh->members.size();
    if (c->rank != root) {
        post_send(c, root, TAG_RED, sbuf, bytes);
        return 0;
    }
    // ascending-rank summation: acc = contribution of rank 0, then += rank 1, ...
    std::vector<char> own(bytes), tmp(bytes), acc(bytes);
    std::memcpy(own.data(), sbuf == MPI_IN_PLACE ? rbuf : sbuf, bytes);
    for (int r = 0; r < n; ++r) {
        const char* src;
        if (r == root) src = own.data();
        else { blocking_recv(c, r, TAG_RED, tmp.data(), bytes); src = tmp.data(); }
        if (r == 0) { std::memcpy(acc.data(), src, bytes); continue; }
        if (dt == MPI_DOUBLE) {
            auto* a = (double*)acc.data(); auto* b = (const double*)src;
            for (int i = 0; i < count; ++i) a[i] += b[i];
        } else if (dt == MPI_INT) {
            auto* a = (int*)acc.data(); auto* b = (const int*)src;
            for (int i = 0; i < count; ++i) a[i] += b[i];
        } else { std::fprintf(stderr, "[mpi_stub] reduce dtype unsupported\n"); std::abort(); }
    }
    std::memcpy(rbuf, acc.data(), bytes);
    return 0;
}
int MPI_Bcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm c) {
    size_t bytes = (size_t)count * dt_size(dt);
    int n = (int)c->sh->members.size();
    if (c->rank == root) {
        for (int r = 0; r < n; ++r) if (r != root) post_send(c, r, TAG_BCAST, buf, bytes);
    } else {
        blocking_recv(c, root, TAG_BCAST, buf, bytes);
    }
    return 0;
}
int MPI_Ibcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm c, MPI_Request* rq) {
    MPI_Bcast(buf, count, dt, root, c);
    *rq = new stub_req; (*rq)->comm = c;
    return 0;
}
int MPI_Allgather(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, int, MPI_Datatype, MPI_Comm c) {
    size_t bytes = (size_t)scount * dt_size(sdt);
    int n = (int)c->sh->members.size();
    for (int r = 0; r < n; ++r) if (r != c->rank) post_send(c, r, TAG_ALLG, sbuf, bytes);
    std::memcpy((char*)rbuf + bytes * c->rank, sbuf, bytes);
    for (int r = 0; r < n; ++r) if (r != c->rank) blocking_recv(c, r, TAG_ALLG, (char*)rbuf + bytes * r, bytes);
    return 0;
}
int MPI_Iscatterv(const void* sbuf, const int* scounts, const int* displs, MPI_Datatype sdt, void* rbuf, int rcount,
                  MPI_Datatype rdt, int root, MPI_Comm c, MPI_Request* rq) {
    int n = (int)c->sh->members.size();
    auto* r = new stub_req; r->comm = c;
    if (c->rank == root) {
        size_t es = dt_size(sdt);
        for (int p = 0; p < n; ++p) {
            const char* src = (const char*)sbuf + (size_t)displs[p] * es;
            size_t bytes = (size_t)scounts[p] * es;
            if (p == root) { if (bytes) std::memmove(rbuf, src, bytes); }
            else post_send(c, p, TAG_SCAT, src, bytes);
        }
    } else {
        r->subs.push_back(SubRecv{rbuf, (size_t)rcount * dt_size(rdt), root, TAG_SCAT, false, 0});
    }
    *rq = r;
    return 0;
}
int MPI_Igatherv(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, const int* rcounts, const int* displs,
                 MPI_Datatype rdt, int root, MPI_Comm c, MPI_Request* rq) {
    int n = (int)c->sh->members.size();
    auto* r = new stub_req; r->comm = c;
    size_t sb = (size_t)scount * dt_size(sdt);
    if (c->rank != root) {
        post_send(c, root, TAG_GATH, sbuf, sb);
    } else {
        size_t es = dt_size(rdt);
        for (int p = 0; p < n; ++p) {
            char* dst = (char*)rbuf + (size_t)displs[p] * es;
            if (p == root) { if (sb) std::memmove(dst, sbuf, sb); }
            else r->subs.push_back(SubRecv{dst, (size_t)rcounts[p] * es, p, TAG_GATH, false, 0});
        }
    }
    *rq = r;
    return 0;
}
int MPI_Igather(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, int rcount, MPI_Datatype rdt, int root,
                MPI_Comm c, MPI_Request* rq) {
    int n = (int)c->sh->members.size();
    std::vector<int> counts(n, rcount), displs(n);
    for (int p = 0; p < n; ++p) displs[p] = p * rcount;
    return MPI_Igatherv(sbuf, scount, sdt, rbuf, counts.data(), displs.data(), rdt, root, c, rq);
}

int MPI_Info_create(MPI_Info* i) { *i = 1; return 0; }
int MPI_Info_set(MPI_Info, const char*, const char*) { return 0; }
int MPI_Info_free(MPI_Info* i) { *i = 0; return 0; }
int MPI_Win_create(void* base, MPI_Aint, int disp_unit, MPI_Info, MPI_Comm c, MPI_Win* w) {
    int seq = c->win_seq++;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto& v = c->sh->winbases[seq];
        if (v.empty()) v.resize(c->sh->members.size(), nullptr);
        v[c->rank] = base;
    }
    barrier(c);
    *w = new stub_win{c->sh, c, seq, disp_unit};
    return 0;
}
int MPI_Win_fence(int, MPI_Win w) { barrier(w->comm); return 0; }
int MPI_Win_free(MPI_Win* w) {
    if (*w) { barrier((*w)->comm); delete *w; *w = MPI_WIN_NULL; }
    return 0;
}
int MPI_Put(const void* origin, int ocount, MPI_Datatype odt, int target, MPI_Aint disp, int, MPI_Datatype,
            MPI_Win w) {
    void* base;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        base = w->sh->winbases[w->seq][target];
    }
    std::memcpy((char*)base + (size_t)disp * w->disp_unit, origin, (size_t)ocount * dt_size(odt));
    return 0;
}

}  // extern "C"
