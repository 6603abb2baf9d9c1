// Decomposition and per-peer block geometry (see geometry.h for the reference citations).
// Pure host code; compiled by nvcc only to keep one toolchain for the library.
#include "geometry.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "kernel_object.h"

namespace dtfftb {

void local_size(int n_global, int comm_dim, int comm_rank, int32_t* start, int32_t* count) {
    // remainder goes to the LAST (n mod p) ranks: src/dtfft_pencil.F90:255-267
    if (comm_dim == 1) {
        *start = 0;
        *count = n_global;
        return;
    }
    const int res = n_global % comm_dim, base = n_global / comm_dim;
    int s = 0;
    for (int q = 0; q < comm_rank; ++q) s += base + (q >= comm_dim - res ? 1 : 0);
    *start = s;
    *count = base + (comm_rank >= comm_dim - res ? 1 : 0);
}

void dims_create(int nnodes, int ndims, int32_t* dims) {
    int fixed = 1, nfree = 0, f0 = -1, f1 = -1;
    for (int i = 0; i < ndims; ++i) {
        if (dims[i] > 0)
            fixed *= dims[i];
        else {
            if (nfree == 0) f0 = i;
            if (nfree == 1) f1 = i;
            ++nfree;
        }
    }
    const int rem = nnodes / fixed;
    if (nfree == 1) {
        dims[f0] = rem;
    } else if (nfree == 2) {
        // balanced factorisation, larger factor first (what MPI_Dims_create returns for two free dims)
        int b = (int)std::floor(std::sqrt((double)rem));
        while (b > 1 && rem % b) --b;
        if (b < 1) b = 1;
        dims[f0] = rem / b;
        dims[f1] = b;
    }
}

GridChoice choose_grid(int ndims, const int32_t* dims, int comm_size, bool cuda, bool z_slab_enabled,
                       bool y_slab_enabled) {
    // src/dtfft_transpose_plan.F90:170-203 (non-cartesian communicator)
    GridChoice g;
    int32_t cd[3] = {1, 0, 0};
    bool cond1 = comm_size <= dims[ndims - 1];
    bool cond2 = comm_size <= dims[0] && comm_size <= dims[1];
    if (cuda) {
        cond1 = kDefTileSize <= dims[ndims - 1] / comm_size;
        cond2 = kDefTileSize <= dims[0] / comm_size && kDefTileSize <= dims[1] / comm_size;
    }
    if (ndims == 3) {
        if (cond1 && z_slab_enabled) {
            cd[1] = 1, cd[2] = comm_size, g.is_z_slab = true;
        } else if (cond2 && y_slab_enabled) {
            cd[1] = comm_size, cd[2] = 1, g.is_y_slab = true;
        } else if (cond1) {
            cd[1] = 1, cd[2] = comm_size;
        } else if (cond2) {
            cd[1] = comm_size, cd[2] = 1;
        }
    }
    dims_create(comm_size, ndims, cd);
    for (int i = 0; i < ndims; ++i) g.comm_dims[i] = cd[i];
    if (dims[ndims - 2] < cd[ndims - 2] || dims[ndims - 1] < cd[ndims - 1]) g.invalid_grid = true;
    return g;
}

std::vector<std::pair<int, int>> grid_candidates(const int32_t* dims, int comm_size) {
    std::vector<std::pair<int, int>> out;
    int dperm[3][3], cperm[3][3];
    permutations(3, dperm, cperm);
    auto valid = [&](int g1, int g2) {
        const int32_t cd[3] = {1, g1, g2};
        for (int d = 0; d < 3; ++d)
            if (dims[dperm[d][1]] < cd[cperm[d][1]] || dims[dperm[d][2]] < cd[cperm[d][2]]) return false;
        return true;
    };
    for (int i = 1; (long long)i * i <= comm_size; ++i) {
        if (comm_size % i) continue;
        const int j = comm_size / i;
        if (valid(i, j)) out.emplace_back(i, j);
        if (i != j && valid(j, i)) out.emplace_back(j, i);
    }
    return out;
}

void cart_coords(int rank, int ndims, const int32_t* comm_dims, int32_t* coords) {
    for (int d = ndims - 1; d >= 0; --d) {
        coords[d] = rank % comm_dims[d];
        rank /= comm_dims[d];
    }
}

int cart_rank(int ndims, const int32_t* comm_dims, const int32_t* coords) {
    int r = 0;
    for (int d = 0; d < ndims; ++d) r = r * comm_dims[d] + coords[d];
    return r;
}

std::vector<int> comm_members(int rank, int ndims, const int32_t* comm_dims, int comm_id) {
    std::vector<int> out;
    int n = 1;
    for (int d = 0; d < ndims; ++d) n *= comm_dims[d];
    if (comm_id == 1) {  // helper%comms(1) = base (cartesian) communicator, abstract_backend.F90:408
        for (int r = 0; r < n; ++r) out.push_back(r);
        return out;
    }
    int32_t coords[3];
    cart_coords(rank, ndims, comm_dims, coords);
    for (int c = 0; c < comm_dims[comm_id - 1]; ++c) {
        int32_t cc[3] = {coords[0], coords[1], coords[2]};
        cc[comm_id - 1] = c;
        out.push_back(cart_rank(ndims, comm_dims, cc));
    }
    return out;
}

void permutations(int ndims, int dperm[3][3], int cperm[3][3]) {
    // src/dtfft_transpose_plan.F90:1046-1082, 0-based
    if (ndims == 2) {
        const int dp[2][2] = {{0, 1}, {1, 0}}, cp[2][2] = {{0, 1}, {0, 1}};
        for (int d = 0; d < 2; ++d)
            for (int j = 0; j < 2; ++j) dperm[d][j] = dp[d][j], cperm[d][j] = cp[d][j];
        return;
    }
    const int dp[3][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}}, cp[3][3] = {{0, 1, 2}, {0, 2, 1}, {0, 1, 2}};
    for (int d = 0; d < 3; ++d)
        for (int j = 0; j < 3; ++j) dperm[d][j] = dp[d][j], cperm[d][j] = cp[d][j];
}

void make_pencils(int ndims, const int32_t* dims, const int32_t* comm_dims, int rank, Pencil out[3]) {
    int dperm[3][3], cperm[3][3];
    permutations(ndims, dperm, cperm);
    int32_t coords[3];
    cart_coords(rank, ndims, comm_dims, coords);
    for (int d = 0; d < ndims; ++d) {
        Pencil& p = out[d];
        p = Pencil{};
        p.aligned_dim = d + 1;
        p.ndims = ndims;
        p.is_even = true;
        for (int j = 0; j < ndims; ++j) {
            const int g = cperm[d][j];
            local_size(dims[dperm[d][j]], comm_dims[g], coords[g], &p.starts[j], &p.counts[j]);
            if (comm_dims[g] > 1 && dims[dperm[d][j]] % comm_dims[g] != 0) p.is_even = false;
        }
        p.is_distributed = false;  // fastest axis is never split in the default decomposition
    }
}

int transpose_comm_id(int ttype) {
    switch (std::abs(ttype)) {
        case 1: return 2;
        case 2: return 3;
        default: return 1;
    }
}

void transpose_pencil_ids(int ttype, int* send, int* recv) {
    switch (ttype) {
        case T_X_TO_Y: *send = 0, *recv = 1; break;
        case T_Y_TO_X: *send = 1, *recv = 0; break;
        case T_Y_TO_Z: *send = 1, *recv = 2; break;
        case T_Z_TO_Y: *send = 2, *recv = 1; break;
        case T_X_TO_Z: *send = 0, *recv = 2; break;
        default: *send = 2, *recv = 0; break;  // Z_TO_X
    }
}

HandleGeometry transpose_geometry(int ttype, const std::vector<Pencil>& send_by_member,
                                  const std::vector<Pencil>& recv_by_member, int me, const std::vector<int>& members,
                                  bool pipelined, bool fused) {
    const int p = (int)members.size();
    const Pencil& send = send_by_member[me];
    const Pencil& recv = recv_by_member[me];
    const int ndims = send.ndims;
    HandleGeometry g;
    g.ttype = ttype;
    g.ndims = ndims;
    g.comm_size = p;
    g.comm_rank = me;
    g.members = members;
    for (int i = 0; i < 3; ++i) g.send_dims[i] = i < ndims ? send.counts[i] : 1, g.recv_dims[i] = i < ndims ? recv.counts[i] : 1;
    const bool forward = ttype == T_X_TO_Y || ttype == T_Y_TO_Z || ttype == T_Z_TO_X;
    int kernel_type = forward ? K_PERMUTE_FORWARD : K_PERMUTE_BACKWARD;  // :232-239
    g.pack_kernel = kernel_type;
    g.has_exchange = p > 1;
    if (!g.has_exchange) return g;  // :246-253

    // 1-based dimension accessors, as in the reference text
    auto S = [&](int d, int r) { return send_by_member[r].counts[d - 1]; };
    auto s = [&](int d, int r) { return send_by_member[r].starts[d - 1]; };
    auto D = [&](int d, int r) { return recv_by_member[r].counts[d - 1]; };
    auto dd = [&](int d, int r) { return recv_by_member[r].starts[d - 1]; };

    // in%ln(:, i), in%ls(:, i) as computed by member `frm` (:295-336)
    auto send_box = [&](int i, int frm, int64_t ln[3], int64_t ls[3]) {
        ln[2] = 1, ls[2] = 0;
        if (ndims == 2) {
            ln[0] = D(2, i), ln[1] = S(2, frm);
            ls[0] = dd(2, i), ls[1] = s(2, frm);
        } else if (ttype == T_X_TO_Z) {
            ln[0] = S(1, frm), ln[1] = D(3, i), ln[2] = S(3, frm);
            ls[0] = s(1, frm), ls[1] = dd(3, i), ls[2] = s(3, frm);
        } else if (ttype == T_Z_TO_X || ttype == T_X_TO_Y || ttype == T_Y_TO_Z) {
            ln[0] = D(3, i), ln[1] = S(2, frm), ln[2] = S(3, frm);
            ls[0] = dd(3, i), ls[1] = s(2, frm), ls[2] = s(3, frm);
        } else {
            ln[0] = D(2, i), ln[1] = S(2, frm), ln[2] = S(3, frm);
            ls[0] = dd(2, i), ls[1] = s(2, frm), ls[2] = s(3, frm);
        }
    };

    g.send_nd.assign(5 * (size_t)p, 0);
    int64_t sdispl = 0;
    for (int i = 0; i < p; ++i) {
        int64_t ln[3], ls[3];
        send_box(i, me, ln, ls);
        int32_t* nd = &g.send_nd[5 * (size_t)i];
        nd[3] = (int32_t)(ttype == T_X_TO_Z ? ln[0] * ls[1] : ls[0]);  // :337-341
        nd[0] = (int32_t)ln[0], nd[1] = (int32_t)ln[1], nd[2] = ndims == 3 ? (int32_t)ln[2] : 1;
        nd[4] = (int32_t)sdispl;  // :413
        const int64_t cnt = ln[0] * ln[1] * (ndims == 3 ? ln[2] : 1);
        g.send_counts.push_back(cnt);
        g.send_displs.push_back(sdispl);
        sdispl += cnt;
    }

    const bool two_step = (ttype == T_Y_TO_X || ttype == T_Z_TO_Y) && ndims == 3 && !fused;  // :430-433
    if (two_step) kernel_type = K_PERMUTE_BACKWARD_START;
    if (fused) {  // get_fused, abstract_kernel.F90:202-217
        if (kernel_type == K_PERMUTE_FORWARD) kernel_type = K_PACK_FORWARD;
        if (kernel_type == K_PERMUTE_BACKWARD) kernel_type = K_PACK_BACKWARD;
    }
    g.pack_kernel = kernel_type;
    g.is_pipelined = pipelined || fused;
    g.is_fused = fused;

    g.recv_nd.assign(5 * (size_t)p, 0);
    int64_t rdispl = 0;
    for (int i = 0; i < p; ++i) {
        int64_t lni[3], lsi[3];
        send_box(me, i, lni, lsi);  // what member i announces it sends to me (:421-422, 488-489)
        const int64_t recvsize = lni[0] * lni[1] * (ndims == 3 ? lni[2] : 1);
        int64_t ln[3] = {0, 0, 0}, ls[3] = {0, 0, 0};
        if (recvsize > 0) {  // :490-531
            if (ndims == 2) {
                ln[0] = S(2, i), ln[1] = D(2, me);
                ls[0] = s(2, i), ls[1] = dd(2, me);
            } else if (ttype == T_X_TO_Z) {
                ln[0] = S(3, i), ln[1] = D(2, me), ln[2] = D(3, me);
                ls[0] = s(3, i), ls[1] = dd(2, me), ls[2] = dd(3, me);
            } else if (ttype == T_Z_TO_X) {
                ln[0] = D(1, me), ln[1] = S(3, i), ln[2] = D(3, me);
                ls[0] = dd(1, me), ls[1] = s(3, i), ls[2] = dd(3, me);
            } else if (ttype == T_X_TO_Y || ttype == T_Y_TO_Z) {
                ln[0] = S(2, i), ln[1] = D(2, me), ln[2] = D(3, me);
                ls[0] = s(2, i), ls[1] = s(2, me), ls[2] = dd(3, me);
            } else {
                ln[0] = S(3, i), ln[1] = D(2, me), ln[2] = D(3, me);
                ls[0] = s(3, i), ls[1] = s(2, me), ls[2] = dd(3, me);
            }
        }
        int32_t* nd = &g.recv_nd[5 * (size_t)i];
        nd[0] = (int32_t)ln[0], nd[1] = (int32_t)ln[1], nd[2] = ndims == 3 ? (int32_t)ln[2] : 1;
        nd[3] = (int32_t)rdispl;                                                 // :582
        nd[4] = (int32_t)(ttype == T_Z_TO_X ? ln[0] * ls[1] : ls[0]);           // :583-588
        g.recv_counts.push_back(recvsize);
        g.recv_displs.push_back(rdispl);
        rdispl += recvsize;
    }

    int uk = K_UNPACK;  // :618-621
    if (g.is_pipelined) uk = K_UNPACK_PIPELINED;
    if (two_step) uk = K_PERMUTE_BACKWARD_END;
    if (g.is_pipelined && two_step) uk = K_PERMUTE_BACKWARD_END_PIPELINED;
    g.unpack_kernel = uk;
    return g;
}

RankLayout layout_of(const Pencil& p) {
    RankLayout l;
    l.ndims = p.ndims;
    int dperm[3][3], cperm[3][3];
    permutations(p.ndims, dperm, cperm);
    for (int j = 0; j < p.ndims; ++j) {
        l.axis[j] = dperm[p.aligned_dim - 1][j];
        l.starts[j] = p.starts[j];
        l.counts[j] = p.counts[j];
    }
    return l;
}

namespace {
struct AxisView {
    long long lo[3], n[3];          // per GLOBAL axis: intersection start / extent
    long long sstride[3], dstride[3];  // per GLOBAL axis: stride in src / dst
    long long sstart[3], dstart[3];
    bool empty = false;
};

// `clip` (optional): per GLOBAL axis [lo, hi) the intersection is further restricted to.
AxisView view_of(const RankLayout& src, const RankLayout& dst, const long long (*clip)[2] = nullptr) {
    AxisView v{};
    const int nd = src.ndims;
    long long st = 1;
    for (int j = 0; j < nd; ++j) {
        v.sstride[src.axis[j]] = st, v.sstart[src.axis[j]] = src.starts[j];
        st *= src.counts[j];
    }
    st = 1;
    for (int j = 0; j < nd; ++j) {
        v.dstride[dst.axis[j]] = st, v.dstart[dst.axis[j]] = dst.starts[j];
        st *= dst.counts[j];
    }
    for (int js = 0; js < nd; ++js) {
        const int A = src.axis[js];
        int jd = 0;
        while (dst.axis[jd] != A) ++jd;
        long long lo = std::max<long long>(src.starts[js], dst.starts[jd]);
        long long hi = std::min<long long>((long long)src.starts[js] + src.counts[js],
                                           (long long)dst.starts[jd] + dst.counts[jd]);
        if (clip) lo = std::max(lo, clip[A][0]), hi = std::min(hi, clip[A][1]);
        v.lo[A] = lo, v.n[A] = hi - lo;
        if (hi <= lo) v.empty = true;
    }
    return v;
}
}  // namespace

namespace {
Box intersect_box_clipped(const RankLayout& src, const RankLayout& dst, const long long (*clip)[2], bool* transposing);
}

Box intersect_box(const RankLayout& src, const RankLayout& dst, bool* transposing) {
    return intersect_box_clipped(src, dst, nullptr, transposing);
}

namespace {
Box intersect_box_clipped(const RankLayout& src, const RankLayout& dst, const long long (*clip)[2], bool* transposing) {
    const int nd = src.ndims;
    const AxisView v = view_of(src, dst, clip);
    const int a = src.axis[0], b = dst.axis[0];
    if (transposing) *transposing = a != b;
    Box x;
    if (v.empty) {
        x.n0 = 0;
        return x;
    }
    for (int j = 0; j < nd; ++j) {
        const int A = src.axis[j];
        x.in_off += (v.lo[A] - v.sstart[A]) * v.sstride[A];
        x.out_off += (v.lo[A] - v.dstart[A]) * v.dstride[A];
    }
    if (a != b) {  // family T: axes (a, b, c)
        int c = -1;
        for (int j = 0; j < nd; ++j)
            if (src.axis[j] != a && src.axis[j] != b) c = src.axis[j];
        x.n0 = v.n[a], x.n1 = v.n[b], x.n2 = c >= 0 ? v.n[c] : 1;
        x.is1 = v.sstride[b], x.is2 = c >= 0 ? v.sstride[c] : 0;
        x.os0 = v.dstride[a], x.os1 = 1, x.os2 = c >= 0 ? v.dstride[c] : 0;
    } else {  // family R: axes in source order
        const int A1 = src.axis[1], A2 = nd == 3 ? src.axis[2] : -1;
        x.n0 = v.n[a], x.n1 = v.n[A1], x.n2 = A2 >= 0 ? v.n[A2] : 1;
        x.is1 = v.sstride[A1], x.is2 = A2 >= 0 ? v.sstride[A2] : 0;
        x.os0 = 1, x.os1 = v.dstride[A1], x.os2 = A2 >= 0 ? v.dstride[A2] : 0;
    }
    return x;
}
}  // namespace

DmaBlock dma_block(const Box& b, bool transposing, long long staging_off) {
    DmaBlock d;
    if (b.empty()) {
        d.ok = true;  // nothing to move
        return d;
    }
    // destination view: a contiguous run and two outer axes (extent, stride)
    long long run, nA, sA, nB, sB;
    if (transposing) {  // family T: output contiguous along b
        run = b.n1, nA = b.n0, sA = b.os0, nB = b.n2, sB = b.os2;
    } else {  // family R: output contiguous along a
        run = b.n0, nA = b.n1, sA = b.os1, nB = b.n2, sB = b.os2;
    }
    bool a_is_rows = true;  // rows = axis A, planes = axis B
    if (nA == 1 && nB > 1) a_is_rows = false;
    if (nA > 1 && nB > 1 && sB < sA) a_is_rows = false;
    const long long n_rows = a_is_rows ? nA : nB, s_rows = a_is_rows ? sA : sB;
    const long long n_planes = a_is_rows ? nB : nA, s_planes = a_is_rows ? sB : sA;
    d.run = run, d.rows = n_rows, d.planes = n_planes;
    d.dst_off = b.out_off;
    d.dst_pitch = n_rows > 1 ? s_rows : run;
    if (n_planes > 1) {
        if (d.dst_pitch <= 0 || s_planes % d.dst_pitch != 0 || s_planes / d.dst_pitch < n_rows) return d;  // ok = false
        d.dst_plane_rows = s_planes / d.dst_pitch;
    } else {
        d.dst_plane_rows = n_rows;
    }
    if (d.dst_pitch < run) return d;
    // pack: same source box, destination = dense [planes][rows][run] at staging_off
    d.pack = b;
    d.pack.out_off = staging_off;
    const long long st_rows = run, st_planes = run * n_rows;
    if (transposing) {
        d.pack.os1 = 1;
        d.pack.os0 = a_is_rows ? st_rows : st_planes;
        d.pack.os2 = a_is_rows ? st_planes : st_rows;
    } else {
        d.pack.os0 = 1;
        d.pack.os1 = a_is_rows ? st_rows : st_planes;
        d.pack.os2 = a_is_rows ? st_planes : st_rows;
    }
    d.ok = true;
    return d;
}

namespace {
// overlap of the two pencils along global axis A
void overlap_along(const RankLayout& a, const RankLayout& b, int A, long long* lo, long long* hi) {
    long long l = 0, h = 1ll << 40;
    for (int j = 0; j < a.ndims; ++j)
        if (a.axis[j] == A) l = std::max<long long>(l, a.starts[j]), h = std::min<long long>(h, (long long)a.starts[j] + a.counts[j]);
    for (int j = 0; j < b.ndims; ++j)
        if (b.axis[j] == A) l = std::max<long long>(l, b.starts[j]), h = std::min<long long>(h, (long long)b.starts[j] + b.counts[j]);
    *lo = l, *hi = std::max(l, h);
}
}  // namespace

int dma_nsub(const Pencil& sender_src, const Pencil& receiver_dst, int64_t base_storage, int group_size) {
    const RankLayout src = layout_of(sender_src), dst = layout_of(receiver_dst);
    bool tr = false;
    const Box b = intersect_box(src, dst, &tr);
    if (b.empty()) return 1;
    // Every copy costs ~13 us of its own on B200 (measured: 7 copies of 33.5 MB in 0.444 ms, 14 in 0.538 ms at 8 GPUs;
    // 1 / 8 copies of 512 / 64 MiB in 0.91 / 0.835 ms at 2 GPUs, profiles/r02e_bench_n8_dma_*.json, r02f_bench_n2_*.json),
    // so a rank issues about 8 copies per exchange: with many peers the peers themselves pipeline packs and copies and the
    // blocks travel whole, with few peers the blocks are cut.  Never below DTFFTB_DMA_SUB_BYTES (default 4 MiB) per slice.
    long long target = 4ll << 20;
    if (const char* e = getenv("DTFFTB_DMA_SUB_BYTES")) target = std::max(1ll, atoll(e));
    long long lo = 0, hi = 0;
    overlap_along(src, dst, src.axis[src.ndims - 1], &lo, &hi);
    long long n = 8 / std::max(1, group_size - 1);
    if (getenv("DTFFTB_DMA_SUB_BYTES")) n = 8;  // explicit slice size: only the cap of 8 applies (tests, A/B runs)
    n = std::min<long long>(n, (b.volume() * base_storage) / target);
    n = std::min<long long>(n, (hi - lo) / 32);
    return (int)std::max<long long>(1, std::min<long long>(8, n));
}

void dma_sub_range(const Pencil& sender_src, const Pencil& receiver_dst, int s, int nsub, int* axis, long long* lo,
                   long long* hi) {
    const RankLayout src = layout_of(sender_src), dst = layout_of(receiver_dst);
    const int A = src.axis[src.ndims - 1];
    long long l = 0, h = 0;
    overlap_along(src, dst, A, &l, &h);
    *axis = A;
    *lo = l + (h - l) * s / nsub;
    *hi = l + (h - l) * (s + 1) / nsub;
}

Box block_box(const Pencil& sender_src, const Pencil& receiver_dst, int s, int nsub, bool* transposing) {
    const RankLayout src = layout_of(sender_src), dst = layout_of(receiver_dst);
    if (nsub <= 1) return intersect_box(src, dst, transposing);
    long long clip[3][2] = {{0, 1ll << 40}, {0, 1ll << 40}, {0, 1ll << 40}};
    int A = 0;
    dma_sub_range(sender_src, receiver_dst, s, nsub, &A, &clip[0][0], &clip[0][1]);
    if (A != 0) {
        clip[A][0] = clip[0][0], clip[A][1] = clip[0][1];
        clip[0][0] = 0, clip[0][1] = 1ll << 40;
    }
    return intersect_box_clipped(src, dst, clip, transposing);
}

Box local_box_for_block(const Pencil& send, const Pencil& recv, const Pencil& x_src, const Pencil& x_dst, int s, int nsub) {
    const RankLayout src = layout_of(send), dst = layout_of(recv), xs = layout_of(x_src), xd = layout_of(x_dst);
    long long clip[3][2] = {{0, 1ll << 40}, {0, 1ll << 40}, {0, 1ll << 40}};
    for (int A = 0; A < src.ndims; ++A) overlap_along(xs, xd, A, &clip[A][0], &clip[A][1]);  // the block's global box
    if (nsub > 1) {
        int A = 0;
        long long lo = 0, hi = 0;
        dma_sub_range(x_src, x_dst, s, nsub, &A, &lo, &hi);
        clip[A][0] = lo, clip[A][1] = hi;
    }
    bool tr = false;
    return intersect_box_clipped(src, dst, clip, &tr);
}

Box local_box_for_peer(const Pencil& send, const Pencil& recv, const Pencil& next_of_peer) {
    const RankLayout src = layout_of(send), dst = layout_of(recv), nxt = layout_of(next_of_peer);
    long long clip[3][2] = {{0, 1ll << 40}, {0, 1ll << 40}, {0, 1ll << 40}};
    for (int j = 0; j < nxt.ndims; ++j) {
        clip[nxt.axis[j]][0] = nxt.starts[j];
        clip[nxt.axis[j]][1] = (long long)nxt.starts[j] + nxt.counts[j];
    }
    bool tr = false;
    return intersect_box_clipped(src, dst, clip, &tr);
}

RankLayout slot_layout(const RankLayout& src, const RankLayout& dst, const RankLayout& order_like) {
    const AxisView v = view_of(src, dst);
    RankLayout s;
    s.ndims = src.ndims;
    for (int j = 0; j < s.ndims; ++j) {
        const int A = order_like.axis[j];
        s.axis[j] = A;
        s.starts[j] = (int32_t)v.lo[A];
        s.counts[j] = v.empty ? 0 : (int32_t)v.n[A];
    }
    return s;
}

std::vector<Box> chunk_boxes(const Pencil& send, const std::vector<Pencil>& recv_by_member, int k, int nchunks,
                             long long* chunk_offset) {
    const int P = (int)recv_by_member.size();
    const int nd = send.ndims;
    const long long n = nd > 0 ? send.counts[nd - 1] : 1;
    const long long lo = n * k / nchunks, hi = n * (k + 1) / nchunks;
    long long slow_stride = 1;  // elements per index of the slowest axis
    for (int j = 0; j + 1 < nd; ++j) slow_stride *= send.counts[j];
    if (chunk_offset) *chunk_offset = lo * slow_stride;
    // the chunk as a layout of its own: same axes, the slowest one restricted to [lo, hi)
    Pencil part = send;
    part.starts[nd - 1] += (int32_t)lo;
    part.counts[nd - 1] = (int32_t)(hi - lo);
    const RankLayout src = layout_of(part);
    std::vector<Box> boxes((size_t)P);
    for (int i = 0; i < P; ++i) {
        bool tr = false;
        boxes[(size_t)i] = hi > lo ? intersect_box(src, layout_of(recv_by_member[(size_t)i]), &tr) : Box{};
    }
    return boxes;
}

namespace {
// pack / unpack boxes and exchange tables of member `m`
void reshape_tables(const std::vector<Pencil>& send_by_member, const std::vector<Pencil>& recv_by_member, int m,
                    ReshapeGeometry* g) {
    const int P = (int)send_by_member.size();
    const RankLayout src = layout_of(send_by_member[(size_t)m]), dst = layout_of(recv_by_member[(size_t)m]);
    g->pack_boxes.assign((size_t)P, Box{});
    g->unpack_boxes.assign((size_t)P, Box{});
    g->send_counts.clear(), g->send_displs.clear(), g->recv_counts.clear(), g->recv_displs.clear();
    int64_t sdispl = 0, rdispl = 0;
    for (int i = 0; i < P; ++i) {
        bool tr = false;
        const RankLayout peer_dst = layout_of(recv_by_member[(size_t)i]);
        const RankLayout slot = slot_layout(src, peer_dst, peer_dst);
        Box pb = intersect_box(src, slot, &tr);
        pb.out_off += sdispl;
        g->pack_boxes[(size_t)i] = pb;
        const int64_t cnt = pb.empty() ? 0 : pb.volume();
        g->send_counts.push_back(cnt), g->send_displs.push_back(sdispl);
        sdispl += cnt;

        const RankLayout peer_src = layout_of(send_by_member[(size_t)i]);
        const RankLayout rslot = slot_layout(peer_src, dst, dst);
        Box ub = intersect_box(rslot, dst, &tr);
        ub.in_off += rdispl;
        g->unpack_boxes[(size_t)i] = ub;
        const int64_t rcnt = ub.empty() ? 0 : ub.volume();
        g->recv_counts.push_back(rcnt), g->recv_displs.push_back(rdispl);
        rdispl += rcnt;
    }
}

// The box is a dense run that already sits where the exchange expects it.
bool identity_placement(const Box& b) {
    if (b.empty()) return true;
    if (b.in_off != b.out_off || b.os0 != 1) return false;
    if (b.n1 > 1 && (b.is1 != b.os1)) return false;
    if (b.n2 > 1 && (b.is2 != b.os2)) return false;
    return true;
}
}  // namespace

ReshapeGeometry reshape_geometry(int rtype, const std::vector<Pencil>& send_by_member,
                                 const std::vector<Pencil>& recv_by_member, int me) {
    const int P = (int)send_by_member.size();
    const int ndims = send_by_member[(size_t)me].ndims;
    ReshapeGeometry g;
    reshape_tables(send_by_member, recv_by_member, me, &g);
    const bool to_pencils = rtype == R_X_BRICKS_TO_PENCILS || rtype == R_Z_BRICKS_TO_PENCILS;

    // reshape_strat of `me` (:267-289)
    if (ndims == 2) {
        g.reshape_strat = 1;
    } else {
        bool zslab = true, yslab = true;
        for (int i = 0; i < P; ++i) {
            if (to_pencils) {
                zslab = zslab && send_by_member[(size_t)me].counts[1] == recv_by_member[(size_t)i].counts[1];
                yslab = yslab && send_by_member[(size_t)me].counts[2] == recv_by_member[(size_t)i].counts[2];
            } else {
                zslab = zslab && send_by_member[(size_t)i].counts[1] == recv_by_member[(size_t)me].counts[1];
                yslab = yslab && send_by_member[(size_t)i].counts[2] == recv_by_member[(size_t)me].counts[2];
            }
        }
        g.reshape_strat = zslab ? 1 : (yslab ? 2 : 3);
    }

    // is_pack_free / is_unpack_free: predicate of every member (:261-266, 479-484) + identity check
    bool ref_all = true, pack_identity = true, unpack_identity = true;
    for (int m = 0; m < P; ++m) {
        bool ok = ndims == 2;
        if (!ok) {
            ok = true;
            for (int i = 0; i < P; ++i)
                ok = ok && send_by_member[(size_t)m].counts[1] == recv_by_member[(size_t)i].counts[1];
        }
        ref_all = ref_all && ok;
        ReshapeGeometry gm;
        const ReshapeGeometry* t = &g;
        if (m != me) {
            reshape_tables(send_by_member, recv_by_member, m, &gm);
            t = &gm;
        }
        for (int i = 0; i < P; ++i) {
            pack_identity = pack_identity && identity_placement(t->pack_boxes[(size_t)i]);
            unpack_identity = unpack_identity && identity_placement(t->unpack_boxes[(size_t)i]);
        }
    }
    g.is_pack_free = to_pencils && ref_all && pack_identity;
    g.is_unpack_free = !to_pencils && ref_all && unpack_identity;
    return g;
}

}  // namespace dtfftb
