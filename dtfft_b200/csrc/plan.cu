// Plan object (see plan.h).  Every block cites the reference lines whose behaviour it restates.
#include "plan.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "errors.h"
#include "trace.h"

namespace dtfftb {

namespace {

bool valid_r2r_kind(int k) { return k >= 3 && k <= 10; }

void split(int n, int p, int r, int32_t* start, int32_t* count) { local_size(n, p, r, start, count); }

}  // namespace

// ======================================================================================
// Plan: creation
// ======================================================================================
void Plan::log(const char* fmt, ...) const {
    if (!cfg_.enable_log || comm_.rank() != 0) return;
    va_list ap;
    va_start(ap, fmt);
    fprintf(stdout, "dtFFT[b200]: ");
    vfprintf(stdout, fmt, ap);
    fprintf(stdout, "\n");
    fflush(stdout);
    va_end(ap);
}

int Plan::create(PlanKind kind, int ndims, const int32_t* dims, const dtfft_pencil_t* pencil, const int* r2r_kinds,
                 const dtfftb_comm_t* comm, int precision, int effort, int executor, bool dry) {
    if (created_) return DTFFT_ERROR_PLAN_IS_CREATED;
    TraceRange trace_create("dtfft_create", kColorCreate);
    dry_ = dry;
    // ---- check_create_args, src/dtfft_plan.F90:2107-2219 ----
    cfg_ = effective_config();
    if (cfg_.platform != DTFFT_PLATFORM_CUDA) return DTFFT_ERROR_INVALID_PLATFORM;
    kind_ = kind;
    comm_ = Comm(comm);
    if (dims) {
        if (ndims != 2 && ndims != 3) return DTFFT_ERROR_INVALID_N_DIMENSIONS;
        for (int i = 0; i < ndims; ++i)
            if (dims[i] <= 0) return DTFFT_ERROR_INVALID_DIMENSION_SIZE;
        ndims_ = ndims;
        for (int i = 0; i < 3; ++i) user_dims_[i] = i < ndims ? dims[i] : 1;
    } else {
        if (!pencil || pencil->ndims == 0) return DTFFT_ERROR_PENCIL_NOT_INITIALIZED;
        ndims_ = pencil->ndims;
    }
    if (precision != DTFFT_SINGLE && precision != DTFFT_DOUBLE) return DTFFT_ERROR_INVALID_PRECISION;
    if (effort < DTFFT_ESTIMATE || effort > DTFFT_EXHAUSTIVE) return DTFFT_ERROR_INVALID_EFFORT;
    if (executor < DTFFT_EXECUTOR_NONE || executor > DTFFT_EXECUTOR_VKFFT) return DTFFT_ERROR_INVALID_EXECUTOR;
    if (executor != DTFFT_EXECUTOR_NONE && executor != DTFFT_EXECUTOR_CUFFT) return DTFFT_ERROR_INVALID_PLATFORM_EXECUTOR;
    precision_ = precision, effort_ = effort, executor_ = executor;
    if (const char* e = getenv("DTFFTB_OVERLAP_CHUNKS")) overlap_chunks_ = std::max(1, atoi(e)), overlap_user_set_ = true;
    if (const char* e = getenv("DTFFTB_OVERLAP_CTAS")) overlap_ctas_ = std::max(0, atoi(e));
    if (const char* e = getenv("DTFFTB_PAIR_OVERLAP")) pair_overlap_ = atoi(e) != 0;
    if (const char* e = getenv("DTFFTB_GRAPHS")) graphs_enabled_ = atoi(e) != 0;
    is_transpose_plan_ = executor == DTFFT_EXECUTOR_NONE;
    if (kind == PLAN_R2R) {
        if (!is_transpose_plan_) {
            if (!r2r_kinds) return DTFFT_ERROR_MISSING_R2R_KINDS;
            for (int i = 0; i < ndims_; ++i)
                if (!valid_r2r_kind(r2r_kinds[i])) return DTFFT_ERROR_INVALID_R2R_KINDS;
            return DTFFT_ERROR_R2R_FFT_NOT_SUPPORTED;  // cuFFT has no r2r (dtfft_executor_cufft_m.F90:94-98)
        }
        if (r2r_kinds)
            for (int i = 0; i < ndims_; ++i) r2r_kinds_[i] = r2r_kinds[i];
    }
    if (kind == PLAN_R2C && is_transpose_plan_) return DTFFT_ERROR_R2C_TRANSPOSE_PLAN;
    // storage sizes, src/dtfft_parameters.F90:301-308
    if (kind == PLAN_R2R)
        base_storage_ = precision == DTFFT_SINGLE ? 4 : 8;
    else
        base_storage_ = precision == DTFFT_SINGLE ? 8 : 16;
    base_storage_init_ = kind == PLAN_R2C ? base_storage_ / 2 : base_storage_;

    if (dry_) {  // host metadata only (decomposition, sizes, geometry); never executes
        int rc0 = choose_decomposition(pencil);
        if (rc0) return rc0;
        rc0 = build_pencils();
        if (rc0) return rc0;
        int wanted0 = cfg_.backend == BACKEND_NONE ? BACKEND_NCCL : cfg_.backend;
        backend_ = comm_.size() == 1 ? BACKEND_NONE : wanted0;
        reshape_backend_ = backend_;
        if (is_final_reshape_enabled_ && is_z_slab_) is_final_reshape_enabled_ = false;
        created_ = true;
        return DTFFT_SUCCESS;
    }
    // ---- device sanity (src/dtfft_plan.F90:1982-2009): one GPU per rank of this host ----
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        return DTFFT_ERROR_GPU_NOT_SET;
    }
    if (cfg_.stream) {
        stream_ = static_cast<cudaStream_t>(cfg_.stream);
        own_stream_ = false;
    } else {  // get_conf_stream, src/dtfft_config.F90:841-852
        ce = cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking);
        if (ce != cudaSuccess) return cuda_error(ce);
        own_stream_ = true;
    }

    int rc = choose_decomposition(pencil);
    if (rc) return rc;
    rc = build_pencils();
    if (rc) return rc;

    const int P = comm_.size();
    // ---- backend choice (src/dtfft_config.F90:766-786; transpose_plan.F90:222) ----
    rc = peers_.init(comm_);
    if (rc) return rc;
    // No backend named by the caller: the reference takes NCCL on CUDA (src/dtfft_config.F90:766-786).  Here the
    // direct-store backend is the default whenever every rank can reach every other rank's memory (one NVSwitch
    // box) and fused backends are enabled -- it takes any device pointer (handle.h: peer_bases) and falls back to
    // NCCL per call for memory cudaIpc cannot share.  DTFFTB_DEFAULT_BACKEND=nccl keeps the reference's choice.
    int wanted = cfg_.backend;
    if (wanted == BACKEND_NONE) {
        const char* e = getenv("DTFFTB_DEFAULT_BACKEND");
        const bool keep_nccl = e && (e[0] == 'n' || e[0] == 'N') && (e[1] == 'c' || e[1] == 'C');
        wanted = (P > 1 && peers_.available() && cfg_.enable_fused_backends && !keep_nccl) ? BACKEND_NVLINK_FUSED : BACKEND_NCCL;
    }
    if (wanted != BACKEND_NCCL && wanted != BACKEND_NCCL_PIPELINED && wanted != BACKEND_NVLINK_FUSED)
        return DTFFT_ERROR_INVALID_BACKEND;
    if (wanted == BACKEND_NVLINK_FUSED && !peers_.available()) {
        log("NVLink peer access unavailable (%s): falling back to the NCCL backend", peers_.why_unavailable());
        wanted = BACKEND_NCCL;
    }
    backend_ = P == 1 ? BACKEND_NONE : wanted;
    // test mode (several ranks time-slicing ONE device, fused backend only): NCCL refuses duplicate devices
    const bool shared_device = peers_.shared_device();
    if (shared_device && wanted != BACKEND_NVLINK_FUSED) return DTFFT_ERROR_GPU_NOT_SET;
    if (P > 1 && !shared_device) {
        rc = init_nccl();
        if (rc) return rc;
    }
    if (P > 1 && effort_ >= DTFFT_PATIENT && backend_candidates().empty())
        return DTFFT_ERROR_BACKENDS_DISABLED;  // reshape_plan_base.F90:139-142
    // grid search applies to default 3-D decompositions only (transpose_plan.F90:230-262)
    bool grid_search = P > 1 && ndims_ == 3 && !has_user_pencil_ && !is_z_slab_ && !is_y_slab_ &&
                       !(comm_.raw() && comm_.raw()->cart_ndims > 0) && effort_ >= DTFFT_MEASURE;
    if (const char* e = getenv("DTFFTB_GRID_SEARCH")) grid_search = grid_search && atoi(e) != 0;
    if (grid_search) {
        rc = autotune_grid(effort_ >= DTFFT_PATIENT);
        if (rc) return rc;
    } else if (P > 1 && effort_ >= DTFFT_PATIENT) {  // run_autotune_backend, transpose_plan.F90:634-855
        rc = autotune_backend();
        if (rc) return rc;
    }
    rc = build_handles(backend_, handles_);
    if (rc) return rc;
    reshape_backend_ = backend_;
    if (cfg_.reshape_backend != BACKEND_NONE && P > 1) {
        reshape_backend_ = cfg_.reshape_backend;
        if (reshape_backend_ == BACKEND_NVLINK_FUSED && !peers_.available()) reshape_backend_ = BACKEND_NCCL;
    }
    if (is_reshape_enabled_) {
        if (P > 1 && effort_ >= DTFFT_EXHAUSTIVE && cfg_.reshape_backend == BACKEND_NONE) {
            rc = autotune_reshape_backend();
            if (rc) return rc;
        }
        rc = build_reshape_handles(reshape_backend_);
        if (rc) return rc;
    }
    if (is_final_reshape_enabled_ && is_z_slab_) is_final_reshape_enabled_ = false;  // dtfft_plan.F90:2093
    if (!is_transpose_plan_) {
        rc = create_ffts();
        if (rc) return rc;
    }
    is_aux_alloc_ = false;
    created_ = true;
    rc = choose_overlap();
    if (rc) {
        created_ = false;
        return rc;
    }
    log("plan created: %dD %s, grid %dx%dx%d, backend %s%s%s", ndims_,
        kind_ == PLAN_C2C ? "c2c" : kind_ == PLAN_R2C ? "r2c" : "r2r", comm_dims_[0], comm_dims_[1], comm_dims_[2],
        dtfft_get_backend_string((dtfft_backend_t)backend_), is_z_slab_ ? ", Z-slab" : "", is_y_slab_ ? ", Y-slab" : "");
    return DTFFT_SUCCESS;
}

// Process grid, global dims and (for user pencils) every rank's X-aligned box.
int Plan::choose_decomposition(const dtfft_pencil_t* pencil) {
    const int P = comm_.size(), me = comm_.rank();
    const int nd = ndims_;
    has_user_pencil_ = pencil != nullptr;
    is_reshape_enabled_ = false;
    is_final_reshape_enabled_ = false;
    is_z_slab_ = is_y_slab_ = false;
    coords_.assign((size_t)P, {0, 0, 0});

    if (!pencil) {
        for (int i = 0; i < 3; ++i) dims_[i] = user_dims_[i];
        if (kind_ == PLAN_R2C) dims_[0] = user_dims_[0] / 2 + 1;  // dtfft_plan.F90:2626
        for (int i = 0; i < 3; ++i) comm_dims_[i] = 1;
        const dtfftb_comm_t* raw = comm_.raw();
        if (raw && raw->cart_ndims > 0) {  // user process grid, transpose_plan.F90:128-170
            const int g = raw->cart_ndims;
            if (g > nd) return DTFFT_ERROR_INVALID_COMM_DIMS;
            long long prod = 1;
            for (int i = 0; i < g; ++i) prod *= raw->cart_dims[i];
            if (prod != P) return DTFFT_ERROR_INVALID_COMM_DIMS;
            if (g == nd) {
                if (raw->cart_dims[0] != 1) return DTFFT_ERROR_INVALID_COMM_FAST_DIM;
                for (int i = 0; i < nd; ++i) comm_dims_[i] = raw->cart_dims[i];
            } else if (g == nd - 1) {
                for (int i = 0; i < g; ++i) comm_dims_[i + 1] = raw->cart_dims[i];
            } else {
                comm_dims_[2] = raw->cart_dims[0];
            }
            if (nd == 3) {
                if (comm_dims_[1] == 1 && cfg_.enable_z_slab)
                    is_z_slab_ = true;
                else if (comm_dims_[2] == 1 && cfg_.enable_y_slab)
                    is_y_slab_ = true;
            }
        } else {  // transpose_plan.F90:171-203
            GridChoice g = choose_grid(nd, dims_, P, true, cfg_.enable_z_slab, cfg_.enable_y_slab);
            for (int i = 0; i < nd; ++i) comm_dims_[i] = g.comm_dims[i];
            is_z_slab_ = g.is_z_slab, is_y_slab_ = g.is_y_slab;
            if (g.invalid_grid) log("WARNING: unable to create correct grid decomposition");
        }
        for (int r = 0; r < P; ++r) {
            int32_t c[3] = {0, 0, 0};
            cart_coords(r, nd, comm_dims_, c);
            coords_[(size_t)r] = {c[0], c[1], c[2]};
        }
        return DTFFT_SUCCESS;
    }

    // ---- pencil_init%create, src/dtfft_pencil.F90:777-908 ----
    int err = DTFFT_SUCCESS;
    if (pencil->ndims < 2 || pencil->ndims > 3) err = DTFFT_ERROR_PENCIL_ARRAYS_INVALID_SIZES;
    if (!err)
        for (int i = 0; i < nd; ++i)
            if (pencil->starts[i] < 0) err = DTFFT_ERROR_PENCIL_INVALID_STARTS;
    if (!err)
        for (int i = 0; i < nd; ++i)
            if (pencil->counts[i] < 0) err = DTFFT_ERROR_PENCIL_INVALID_COUNTS;
    err = (int)comm_.max((double)err);  // CHECK_ERROR_AND_RETURN_AGG
    if (err) return err;
    UserPencil mine{};
    for (int i = 0; i < 3; ++i) mine.starts[i] = i < nd ? pencil->starts[i] : 0, mine.counts[i] = i < nd ? pencil->counts[i] : 1;
    if (comm_.allgather_v(mine, user_pencils_)) return DTFFTB_ERROR_COMM;
    auto& up = user_pencils_;
    for (int d = 0; d < 3; ++d) user_dims_[d] = 1;
    for (int d = 0; d < nd; ++d)
        for (int r = 0; r < P; ++r) user_dims_[d] = std::max(user_dims_[d], up[r].starts[d] + up[r].counts[d]);
    auto empty = [&](int r) {
        for (int d = 0; d < nd; ++d)
            if (up[r].counts[d] == 0) return true;
        return false;
    };
    for (int p1 = 0; p1 < P && !err; ++p1) {  // :838-862 shape mismatch
        if (empty(p1)) continue;
        for (int p2 = p1 + 1; p2 < P; ++p2) {
            if (empty(p2)) continue;
            for (int d1 = 0; d1 < nd; ++d1)
                for (int d2 = d1 + 1; d2 < nd; ++d2)
                    if (up[p1].starts[d1] == up[p2].starts[d1] && up[p1].starts[d2] == up[p2].starts[d2] &&
                        (up[p1].counts[d1] != up[p2].counts[d1] || up[p1].counts[d2] != up[p2].counts[d2]))
                        err = DTFFT_ERROR_PENCIL_SHAPE_MISMATCH;
        }
    }
    if (err) return err;
    for (int i = 0; i < P && !err; ++i)  // :865-873 overlap
        for (int j = i + 1; j < P; ++j) {
            if (empty(i) || empty(j)) continue;
            bool ov = true;
            for (int d = 0; d < nd; ++d)
                if (up[i].starts[d] + up[i].counts[d] <= up[j].starts[d] || up[j].starts[d] + up[j].counts[d] <= up[i].starts[d])
                    ov = false;
            if (ov) err = DTFFT_ERROR_PENCIL_OVERLAP;
        }
    if (err) return err;
    {  // :875-879 continuity
        long long vol = 0, gvol = 1;
        for (int r = 0; r < P; ++r) {
            if (empty(r)) continue;
            long long v = 1;
            for (int d = 0; d < nd; ++d) v *= up[r].counts[d];
            vol += v;
        }
        for (int d = 0; d < nd; ++d) gvol *= user_dims_[d];
        if (vol != gvol) return DTFFT_ERROR_PENCIL_NOT_CONTINUOUS;
    }
    // 1-D communicators of the user's grid (create_1d_comm, :1034-1081): ranks that share
    // start and extent on every other axis, ordered by their start along the axis
    int32_t bgrid[3] = {1, 1, 1};
    std::vector<std::array<int32_t, 3>> bcoord((size_t)P, {0, 0, 0});
    for (int r = 0; r < P; ++r)
        for (int d = 0; d < nd; ++d) {
            // a rank with no points along d may share its start with a neighbour (:1066-1068);
            // ties keep rank order (stable insertion sort, :1005-1031)
            std::vector<std::pair<int32_t, int>> line;
            for (int i = 0; i < P; ++i) {
                bool same = true;
                for (int j = 0; j < nd; ++j)
                    if (j != d && (up[i].starts[j] != up[r].starts[j] || up[i].counts[j] != up[r].counts[j])) same = false;
                const bool tie_with_empty = up[i].starts[d] == up[r].starts[d] && (up[i].counts[d] == 0 || up[r].counts[d] == 0);
                if (i == r || (same && (up[i].starts[d] != up[r].starts[d] || tie_with_empty))) line.push_back({up[i].starts[d], i});
            }
            std::sort(line.begin(), line.end());
            int idx = (int)(std::lower_bound(line.begin(), line.end(), std::make_pair(up[r].starts[d], r)) - line.begin());
            bcoord[(size_t)r][(size_t)d] = idx;
            if (r == me) bgrid[d] = (int32_t)line.size();
        }
    {
        long long prod = 1;
        for (int d = 0; d < nd; ++d) prod *= bgrid[d];
        int bad = prod != P;
        if (comm_.max(bad) > 0) return DTFFT_ERROR_PENCIL_NOT_CONTINUOUS;
    }
    for (int i = 0; i < 3; ++i) dims_[i] = user_dims_[i];
    if (kind_ == PLAN_R2C) dims_[0] = user_dims_[0] / 2 + 1;

    if (bgrid[0] == 1) {  // X pencils supplied: keep the user's grid (dtfft_plan.F90:2063-2074)
        for (int d = 0; d < 3; ++d) comm_dims_[d] = d < nd ? bgrid[d] : 1;
        coords_ = bcoord;
    } else {
        // ---- from_bricks, src/dtfft_pencil.F90:520-775 ----
        is_reshape_enabled_ = true;
        const int fast = bgrid[0];
        const int tile = kDefTileSize;
        int y_size = 1, z_size = 1;
        bool nice = false;
        const UserPencil& b0 = up[0];  // rank 0 decides (MPI_Bcast, :640-645)
        if (nd == 3 && b0.counts[2] > tile * fast) {
            y_size = 1, z_size = fast, nice = true;
        } else if (b0.counts[1] > tile * fast || nd == 2) {
            y_size = fast, z_size = 1, nice = true;
        } else {
            for (int i = 2; i <= fast; ++i) {
                if (fast % i) continue;
                if (b0.counts[2] < tile * i || b0.counts[2] % i) continue;
                if (b0.counts[1] < tile * i || b0.counts[1] % i) continue;
                nice = true;
                y_size = std::min(i, fast / i), z_size = std::max(i, fast / i);
                break;
            }
        }
        if (!nice) {
            int32_t t[2] = {0, 0};
            dims_create(fast, 2, t);
            y_size = t[0], z_size = t[1];
            log("WARNING: unable to find good grid decomposition, using MPI_Dims_create");
        }
        comm_dims_[0] = 1, comm_dims_[1] = bgrid[1] * y_size, comm_dims_[2] = nd == 3 ? bgrid[2] * z_size : 1;
        bricks_.assign((size_t)P, {});
        xpencil_from_bricks_.assign((size_t)P, UserPencil{});
        for (int r = 0; r < P; ++r) {
            const int a = bcoord[(size_t)r][0], b = bcoord[(size_t)r][1], c = nd == 3 ? bcoord[(size_t)r][2] : 0;
            int ay = 0, az = 0;
            if (y_size == 1)
                az = a;
            else if (z_size == 1)
                ay = a;
            else
                ay = a / z_size, az = a % z_size;
            coords_[(size_t)r] = {0, b * y_size + ay, c * z_size + az};
            UserPencil x{};
            x.starts[0] = 0, x.counts[0] = user_dims_[0];
            int32_t s, n;
            split(up[r].counts[1], y_size, ay, &s, &n);
            x.starts[1] = up[r].starts[1] + s, x.counts[1] = n;
            if (nd == 3) {
                split(up[r].counts[2], z_size, az, &s, &n);
                x.starts[2] = up[r].starts[2] + s, x.counts[2] = n;
            } else {
                x.starts[2] = 0, x.counts[2] = 1;
            }
            xpencil_from_bricks_[(size_t)r] = x;
            Pencil bp;
            bp.aligned_dim = 1, bp.ndims = nd;
            for (int d = 0; d < nd; ++d) bp.starts[d] = up[r].starts[d], bp.counts[d] = up[r].counts[d];
            bricks_[(size_t)r][0] = bp;
        }
        brick_grid_[0] = bgrid[0], brick_grid_[1] = bgrid[1], brick_grid_[2] = bgrid[2];
        brick_coords_ = bcoord;
    }
    if (nd == 3) {  // transpose_plan.F90:113-125
        if (comm_dims_[1] == 1 && cfg_.enable_z_slab)
            is_z_slab_ = true;
        else if (comm_dims_[2] == 1 && cfg_.enable_y_slab)
            is_y_slab_ = true;
    }
    return DTFFT_SUCCESS;
}

// X / Y / Z pencils of every rank (create_pencils_and_comm, transpose_plan.F90:1084-1131, with
// the carry-over rule of pencil%create, src/dtfft_pencil.F90:136-163, for user pencils).
int Plan::build_pencils() {
    const int P = comm_.size(), nd = ndims_;
    pencils_.assign((size_t)P, {});
    real_pencils_.assign((size_t)P, Pencil{});
    for (int r = 0; r < P; ++r) {
        const auto& c = coords_[(size_t)r];
        Pencil X, Y, Z;
        X.aligned_dim = 1, Y.aligned_dim = 2, Z.aligned_dim = 3;
        X.ndims = Y.ndims = Z.ndims = nd;
        X.starts[0] = 0, X.counts[0] = dims_[0];
        if (has_user_pencil_) {
            const UserPencil& u = is_reshape_enabled_ ? xpencil_from_bricks_[(size_t)r] : user_pencils_[(size_t)r];
            for (int d = 1; d < nd; ++d) X.starts[d] = u.starts[d], X.counts[d] = u.counts[d];
        } else {
            for (int d = 1; d < nd; ++d) split(dims_[d], comm_dims_[d], c[(size_t)d], &X.starts[d], &X.counts[d]);
        }
        if (nd == 2) {
            Y.starts[0] = 0, Y.counts[0] = dims_[1];
            split(dims_[0], comm_dims_[1], c[1], &Y.starts[1], &Y.counts[1]);
        } else {
            Y.starts[0] = 0, Y.counts[0] = dims_[1];
            Y.starts[1] = X.starts[2], Y.counts[1] = X.counts[2];  // z keeps its split
            split(dims_[0], comm_dims_[1], c[1], &Y.starts[2], &Y.counts[2]);
            Z.starts[0] = 0, Z.counts[0] = dims_[2];
            Z.starts[1] = Y.starts[2], Z.counts[1] = Y.counts[2];  // x keeps its split
            split(dims_[1], comm_dims_[2], c[2], &Z.starts[2], &Z.counts[2]);
        }
        pencils_[(size_t)r] = {X, Y, Z};
        Pencil R = X;  // real-space X pencil of an R2C plan (dtfft_plan.F90:2630)
        R.counts[0] = user_dims_[0];
        real_pencils_[(size_t)r] = R;
    }
    if (is_reshape_enabled_) {
        // Z bricks (reshape_plan.F90:150-182): groups of `c` consecutive ranks of the last grid
        // dimension split z among themselves and pool their share of the slowest axis
        const int last = nd - 1;
        const int csize = brick_grid_[last];
        const int gsize = comm_dims_[last];
        for (int r = 0; r < P; ++r) {
            const auto& c = coords_[(size_t)r];
            const Pencil& L = pencils_[(size_t)r][(size_t)last];  // last pencil: (z,x,y) or (y,x)
            const int grp = c[(size_t)last] / csize, k = c[(size_t)last] % csize;
            Pencil B;
            B.aligned_dim = nd, B.ndims = nd;
            split(dims_[last], csize, k, &B.starts[0], &B.counts[0]);
            // pooled slowest axis over the group
            int lo = 1 << 30, cnt = 0;
            for (int q = 0; q < P; ++q) {
                const auto& cq = coords_[(size_t)q];
                bool same = true;
                for (int d = 1; d < nd; ++d)
                    if (d != last && cq[(size_t)d] != c[(size_t)d]) same = false;
                if (!same || cq[(size_t)last] / csize != grp) continue;
                const Pencil& Lq = pencils_[(size_t)q][(size_t)last];
                lo = std::min(lo, (int)Lq.starts[nd - 1]);
                cnt += Lq.counts[nd - 1];
            }
            if (nd == 3) B.starts[1] = L.starts[1], B.counts[1] = L.counts[1];
            B.starts[nd - 1] = lo, B.counts[nd - 1] = cnt;
            bricks_[(size_t)r][1] = B;
        }
        (void)gsize;
        is_final_reshape_enabled_ = csize > 1 && cfg_.enable_fourier_reshape;  // reshape_plan.F90:190-191
    }
    return DTFFT_SUCCESS;
}

int Plan::init_nccl() {
    // backend_helper%create, src/dtfft_abstract_backend.F90:432-457 (MPI_Bcast -> allgather)
    if (nccl_) return DTFFT_SUCCESS;
    ncclUniqueId id;
    std::memset(&id, 0, sizeof(id));
    if (comm_.rank() == 0) {
        ncclResult_t nr = ncclGetUniqueId(&id);
        if (nr != ncclSuccess) return nccl_error(nr);
    }
    std::vector<ncclUniqueId> all;
    if (comm_.allgather_v(id, all)) return DTFFTB_ERROR_COMM;
    ncclResult_t nr = ncclCommInitRank(&nccl_, comm_.size(), all[0], comm_.rank());
    if (nr != ncclSuccess) return nccl_error(nr);
    return DTFFT_SUCCESS;
}

std::vector<int> Plan::group_members(int rank, int comm_id) const {
    const int P = comm_.size(), nd = ndims_;
    std::vector<std::pair<long long, int>> keyed;
    const auto& c = coords_[(size_t)rank];
    for (int r = 0; r < P; ++r) {
        const auto& q = coords_[(size_t)r];
        if (comm_id == 1) {
            keyed.push_back({(long long)q[1] * comm_dims_[2] + q[2], r});
        } else {
            bool same = true;
            for (int d = 1; d < nd; ++d)
                if (d != comm_id - 1 && q[(size_t)d] != c[(size_t)d]) same = false;
            if (same) keyed.push_back({q[(size_t)(comm_id - 1)], r});
        }
    }
    std::sort(keyed.begin(), keyed.end());
    std::vector<int> out;
    for (auto& k : keyed) out.push_back(k.second);
    return out;
}

int Plan::handle_spec(int type, HandleSpec* hs) const {
    const int P = comm_.size(), me = comm_.rank(), nd = ndims_;
    hs->members.clear(), hs->send.clear(), hs->recv.clear();
    if (std::abs(type) <= 3) {  // transposition: plans(-3:3), transpose_plan.F90:334-361
        const int at = std::abs(type);
        if (at < 1 || (nd == 2 && at > 1) || (at == 3 && !is_z_slab_)) return DTFFT_ERROR_INVALID_TRANSPOSE_TYPE;
        hs->ttype = type, hs->rtype = 0;
        hs->comm_id = transpose_comm_id(type);
        hs->members = group_members(me, hs->comm_id);
        int si, ri;
        transpose_pencil_ids(type, &si, &ri);
        for (int m : hs->members)
            hs->send.push_back(pencils_[(size_t)m][(size_t)si]), hs->recv.push_back(pencils_[(size_t)m][(size_t)ri]);
        hs->es = base_storage_;
    } else {  // reshape: plans(11:14), reshape_plan.F90:405-438; X reshapes move real elements for R2C
        if (!is_reshape_enabled_) return DTFFT_ERROR_RESHAPE_NOT_SUPPORTED;
        if (type < R_X_BRICKS_TO_PENCILS || type > R_Z_BRICKS_TO_PENCILS) return DTFFT_ERROR_INVALID_RESHAPE_TYPE;
        hs->ttype = 0, hs->rtype = type;
        const bool x_side = type == R_X_BRICKS_TO_PENCILS || type == R_X_PENCILS_TO_BRICKS;
        const bool to_pencils = type == R_X_BRICKS_TO_PENCILS || type == R_Z_BRICKS_TO_PENCILS;
        const int last = nd - 1, csize = brick_grid_[last];
        std::vector<std::pair<std::pair<long long, int>, int>> keyed;
        for (int r = 0; r < P; ++r) {
            bool same = true;
            if (x_side) {  // bricks sharing their (y, z) footprint = one line of the brick grid along x,
                           // ranked by where their X pencil starts (MPI_Comm_split key, reshape_plan.F90:162-167)
                for (int d = 1; d < nd; ++d)
                    if (brick_coords_[(size_t)r][(size_t)d] != brick_coords_[(size_t)me][(size_t)d]) same = false;
                const Pencil& xp = pencils_[(size_t)r][0];
                long long key = xp.starts[1];
                if (nd == 3) key += (long long)xp.starts[2] * xp.counts[1] * brick_grid_[1];
                if (same) keyed.push_back({{key, brick_coords_[(size_t)r][0]}, r});
            } else {  // the `c` consecutive ranks of the last grid dimension that pool their data
                for (int d = 1; d < nd; ++d)
                    if (d != last && coords_[(size_t)r][(size_t)d] != coords_[(size_t)me][(size_t)d]) same = false;
                if (same && coords_[(size_t)r][(size_t)last] / csize == coords_[(size_t)me][(size_t)last] / csize)
                    keyed.push_back({{coords_[(size_t)r][(size_t)last], 0}, r});
            }
        }
        std::sort(keyed.begin(), keyed.end());
        for (auto& k : keyed) hs->members.push_back(k.second);
        for (int m : hs->members) {
            const Pencil& brick = bricks_[(size_t)m][x_side ? 0 : 1];
            const Pencil& pen = x_side ? (kind_ == PLAN_R2C ? real_pencils_[(size_t)m] : pencils_[(size_t)m][0])
                                       : pencils_[(size_t)m][(size_t)last];
            hs->send.push_back(to_pencils ? brick : pen);
            hs->recv.push_back(to_pencils ? pen : brick);
        }
        hs->comm_id = x_side ? 4 : 5;  // barrier channels of their own
        hs->es = x_side ? base_storage_init_ : base_storage_;
    }
    hs->me = (int)(std::find(hs->members.begin(), hs->members.end(), me) - hs->members.begin());
    return DTFFT_SUCCESS;
}

std::vector<int> Plan::transpose_types() const {
    std::vector<int> types;
    for (int d = 1; d < ndims_; ++d) types.push_back(d), types.push_back(-d);
    if (is_z_slab_) types.push_back(3), types.push_back(-3);
    return types;
}

int Plan::build_handles(int backend, std::map<int, std::unique_ptr<ReshapeHandle>>& into) {
    into.clear();
    HandleContext ctx;
    ctx.nccl = nccl_;
    ctx.peers = &peers_;
    ctx.effort = cfg_.enable_kernel_autotune ? DTFFT_EXHAUSTIVE : effort_;
    for (int t : transpose_types()) {
        HandleSpec hs;
        int rc = handle_spec(t, &hs);
        if (rc) return rc;
        std::unique_ptr<ReshapeHandle> h(new ReshapeHandle);
        int b = hs.members.size() > 1 ? backend : BACKEND_NONE;
        rc = h->create(ctx, t, 0, hs.comm_id, hs.members, hs.me, hs.send, hs.recv, hs.es, b);
        if (rc) return rc;
        into[t] = std::move(h);
    }
    return DTFFT_SUCCESS;
}

int Plan::build_reshape_handles(int backend) { return build_reshape_handles(backend, rhandles_); }

int Plan::build_reshape_handles(int backend, std::map<int, std::unique_ptr<ReshapeHandle>>& rhandles_) {
    rhandles_.clear();
    HandleContext ctx;
    ctx.nccl = nccl_;
    ctx.peers = &peers_;
    ctx.effort = effort_;
    for (int t = R_X_BRICKS_TO_PENCILS; t <= R_Z_BRICKS_TO_PENCILS; ++t) {
        HandleSpec hs;
        int rc = handle_spec(t, &hs);
        if (rc) return rc;
        std::unique_ptr<ReshapeHandle> h(new ReshapeHandle);
        int b = hs.members.size() > 1 ? backend : BACKEND_NONE;
        rc = h->create(ctx, 0, t, hs.comm_id, hs.members, hs.me, hs.send, hs.recv, hs.es, b);
        if (rc) return rc;
        rhandles_[t] = std::move(h);
    }
    return DTFFT_SUCCESS;
}

// Introspection for tests / report: the exchange geometry of one transposition or reshape.
int Plan::describe_exchange(int type, ExchangeDescription* d) const {
    HandleSpec hs;
    int rc = handle_spec(type, &hs);
    if (rc) return rc;
    const int P = (int)hs.members.size();
    d->members = hs.members;
    d->me = hs.me;
    d->element_bytes = hs.es;
    d->geo = HandleGeometry{};
    if (hs.ttype != 0)
        d->geo = transpose_geometry(hs.ttype, hs.send, hs.recv, hs.me, hs.members, backend_is_pipelined(backend_), false);
    d->fused.assign((size_t)P, Box{});
    d->fused_transposing = false;
    const RankLayout src = layout_of(hs.send[(size_t)hs.me]);
    for (int i = 0; i < P; ++i) {
        bool tr = false;
        d->fused[(size_t)i] = intersect_box(src, layout_of(hs.recv[(size_t)i]), &tr);
        if (!d->fused[(size_t)i].empty() && tr) d->fused_transposing = true;
    }
    return DTFFT_SUCCESS;
}

int Plan::describe_reshape(int rtype, std::vector<int>* members, int* me, ReshapeGeometry* g) const {
    HandleSpec hs;
    if (std::abs(rtype) <= 3) return DTFFT_ERROR_INVALID_RESHAPE_TYPE;
    int rc = handle_spec(rtype, &hs);
    if (rc) return rc;
    if (members) *members = hs.members;
    if (me) *me = hs.me;
    *g = reshape_geometry(rtype, hs.send, hs.recv, hs.me);
    return DTFFT_SUCCESS;
}

int Plan::describe_chunk(int ttype, int k, int nchunks, std::vector<int>* members, std::vector<Box>* boxes,
                         long long* chunk_offset) const {
    if (std::abs(ttype) < 1 || std::abs(ttype) > 3) return DTFFT_ERROR_INVALID_TRANSPOSE_TYPE;
    if (nchunks < 1 || k < 0 || k >= nchunks) return DTFFT_ERROR_INVALID_USAGE;
    HandleSpec hs;
    int rc = handle_spec(ttype, &hs);
    if (rc) return rc;
    if (members) *members = hs.members;
    *boxes = chunk_boxes(hs.send[(size_t)hs.me], hs.recv, k, nchunks, chunk_offset);
    return DTFFT_SUCCESS;
}

int Plan::time_backend(int backend, double* ms, bool reshapes) {
    // execute_autotune, src/dtfft_reshape_plan_base.F90:588-706: every transposition (or every
    // reshape) once per iteration, warm-up + timed iterations, result = max over ranks of the mean
    std::map<int, std::unique_ptr<ReshapeHandle>> hs;
    int rc = reshapes ? build_reshape_handles(backend, hs) : build_handles(backend, hs);
    *ms = 1e30;
    int ok = rc == DTFFT_SUCCESS;
    if (comm_.sum(ok) != comm_.size()) return DTFFT_SUCCESS;  // backend unusable somewhere: skip it
    size_t bytes = alloc_bytes(), aux = 0;
    for (auto& kv : hs) aux = std::max(aux, (size_t)kv.second->aux_bytes());
    const int saved = backend_, saved_r = reshape_backend_;
    backend_ = backend;  // mem_alloc picks the allocator by backend
    if (reshapes) reshape_backend_ = backend;
    void *a = nullptr, *b = nullptr, *w = nullptr;
    rc = mem_alloc(bytes, &a);
    if (!rc) rc = mem_alloc(bytes, &b);
    if (!rc && aux) rc = mem_alloc(aux, &w);
    // every rank must take the same branch: the timed branch ends in a collective max
    const bool all_ok = comm_.sum(rc == DTFFT_SUCCESS ? 1 : 0) == comm_.size();
    if (!all_ok && !rc) rc = DTFFT_ERROR_ALLOC_FAILED;
    if (all_ok) {
        cudaMemsetAsync(a, 0, bytes, stream_);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        auto pass = [&]() {
            int r2 = 0;
            for (auto& kv : hs) {
                r2 = kv.second->execute(a, b, stream_, w);
                if (r2) return r2;
            }
            return r2;
        };
        for (int i = 0; i < cfg_.n_measure_warmup_iters && !rc; ++i) rc = pass();
        cudaEventRecord(e0, stream_);
        for (int i = 0; i < cfg_.n_measure_iters && !rc; ++i) rc = pass();
        cudaEventRecord(e1, stream_);
        cudaError_t ce = cudaEventSynchronize(e1);
        float t = 0;
        if (!rc && ce == cudaSuccess) {
            cudaEventElapsedTime(&t, e0, e1);
            // report_timings, src/dtfft_reshape_plan_base.F90:708-731: max / min / avg over the ranks of the mean
            // time per iteration; the max decides
            const double mine = (double)t / std::max(1, (int)cfg_.n_measure_iters);
            std::vector<double> all;
            comm_.allgather_v(mine, all);
            double mx = all[0], mn = all[0], sum = 0;
            for (double x : all) mx = std::max(mx, x), mn = std::min(mn, x), sum += x;
            *ms = mx;
            log("  %s %s:", reshapes ? "reshapes," : "transpositions,", dtfft_get_backend_string((dtfft_backend_t)backend));
            log("    max: %.6f [ms]", mx);
            log("    min: %.6f [ms]", mn);
            log("    avg: %.6f [ms]", sum / (double)all.size());
        } else {
            comm_.max(1e30);
        }
        cudaEventDestroy(e0), cudaEventDestroy(e1);
    }
    hs.clear();
    if (w) mem_free(w);
    if (b) mem_free(b);
    if (a) mem_free(a);
    backend_ = saved;
    reshape_backend_ = saved_r;
    return rc;
}

std::vector<int> Plan::backend_candidates() const {
    // run_autotune_backend's filter (src/dtfft_transpose_plan.F90:763-781) on the backends that exist here
    std::vector<int> cands;
    if (cfg_.enable_nccl_backends) cands.push_back(BACKEND_NCCL);
    if (cfg_.enable_nccl_backends && cfg_.enable_pipelined_backends) cands.push_back(BACKEND_NCCL_PIPELINED);
    if (cfg_.enable_fused_backends && peers_.available()) cands.push_back(BACKEND_NVLINK_FUSED);
    return cands;
}

void Plan::set_grid(int g1, int g2) {
    comm_dims_[0] = 1, comm_dims_[1] = g1, comm_dims_[2] = g2;
    const int P = comm_.size();
    coords_.assign((size_t)P, {0, 0, 0});
    for (int r = 0; r < P; ++r) {
        int32_t c[3] = {0, 0, 0};
        cart_coords(r, ndims_, comm_dims_, c);
        coords_[(size_t)r] = {c[0], c[1], c[2]};
    }
    build_pencils();
}

int Plan::dry_set_grid(int g1, int g2) {
    if (!created_ || !dry_ || ndims_ != 3 || has_user_pencil_) return DTFFT_ERROR_INVALID_USAGE;
    if (g1 < 1 || g2 < 1 || (long long)g1 * g2 != comm_.size()) return DTFFT_ERROR_INVALID_COMM_DIMS;
    is_z_slab_ = is_y_slab_ = false;
    set_grid(g1, g2);
    return DTFFT_SUCCESS;
}

int Plan::autotune_grid(bool all_backends) {
    TraceRange trace("Autotune transpose plan", kColorAutotune);  // transpose_plan.F90:233
    const int saved1 = comm_dims_[1], saved2 = comm_dims_[2];
    const std::vector<std::pair<int, int>> grids = grid_candidates(dims_, comm_.size());
    const std::vector<int> backends = all_backends ? backend_candidates() : std::vector<int>{backend_};
    double best = 1e30;
    int best_b = backend_, best1 = saved1, best2 = saved2;
    for (const auto& g : grids) {
        set_grid(g.first, g.second);
        int rcb = peers_.reset_barriers();  // the 1-D communicators just changed
        if (rcb) return rcb;
        for (int b : backends) {
            double ms = 1e30;
            int rc = time_backend(b, &ms);
            if (rc) return rc;
            log("autotune grid 1x%dx%d, backend %s: %.4f ms", g.first, g.second, dtfft_get_backend_string((dtfft_backend_t)b), ms);
            if (ms < best) best = ms, best_b = b, best1 = g.first, best2 = g.second;
        }
    }
    // nothing could be timed (no valid grid): keep the default decomposition (transpose_plan.F90:502-525)
    set_grid(best1, best2);
    int rcb = peers_.reset_barriers();
    if (rcb) return rcb;
    backend_ = best_b;
    log("DTFFT_MEASURE: selected process grid 1x%dx%d", best1, best2);
    if (all_backends) log("DTFFT_PATIENT: selected backend is %s", dtfft_get_backend_string((dtfft_backend_t)backend_));
    return DTFFT_SUCCESS;
}

int Plan::autotune_reshape_backend() {
    TraceRange trace("Autotune reshape plan", kColorAutotune);  // reshape_plan.F90:209
    // autotune_reshape_plan (src/dtfft_reshape_plan.F90:206-222, 487-560): the four reshapes timed
    // with every enabled backend, the fastest kept
    const std::vector<int> cands = backend_candidates();
    double best = 1e30;
    int best_b = reshape_backend_;
    for (int b : cands) {
        double ms = 1e30;
        int rc = time_backend(b, &ms, true);
        if (rc) return rc;
        log("autotune reshape backend %s: %.4f ms", dtfft_get_backend_string((dtfft_backend_t)b), ms);
        if (ms < best) best = ms, best_b = b;
    }
    reshape_backend_ = best_b;
    log("DTFFT_EXHAUSTIVE: selected reshape backend is %s", dtfft_get_backend_string((dtfft_backend_t)reshape_backend_));
    return DTFFT_SUCCESS;
}

int Plan::autotune_backend() {
    TraceRange trace("Autotune transpose plan", kColorAutotune);
    const std::vector<int> cands = backend_candidates();
    double best = 1e30;
    int best_b = backend_;
    for (int b : cands) {
        double ms = 1e30;
        int rc = time_backend(b, &ms);
        if (rc) return rc;
        log("autotune backend %s: %.4f ms", dtfft_get_backend_string((dtfft_backend_t)b), ms);
        if (ms < best) best = ms, best_b = b;
    }
    backend_ = best_b;
    log("DTFFT_PATIENT: selected backend is %s", dtfft_get_backend_string((dtfft_backend_t)backend_));
    return DTFFT_SUCCESS;
}

int Plan::choose_overlap() {
    // Stage overlap (FFT chunk k+1 || exchange of chunk k) pays when the FFT before an exchange is
    // long relative to the exchange: measured on 2 and 8 B200 (profiles/r01d_configs_n*.jsonl) it wins
    // ~8 % on the 16384^2 slab (16384-point transforms) and loses on 512^3 / 1024^3 pencils, whose
    // short transforms cannot amortise the extra launches.  Every rank must take the same decision
    // (the backward pencil schedule changes its buffer choreography with it), so the rule only uses
    // plan-wide quantities; effort >= DTFFT_MEASURE replaces the rule by a timed choice.
    const int P = comm_.size();
    const bool applicable = !is_transpose_plan_ && P > 1 && backend_ == BACKEND_NVLINK_FUSED;
    if (overlap_user_set_) {
        if (!applicable) overlap_chunks_ = 1;
        return DTFFT_SUCCESS;
    }
    overlap_chunks_ = 1;
    if (!applicable) return DTFFT_SUCCESS;
    long long longest = 1, total = 1;
    for (int d = 0; d < ndims_; ++d) longest = std::max<long long>(longest, dims_[d]), total *= dims_[d];
    const long long local_bytes = total * base_storage_ / P;
    if (longest >= 4096 && local_bytes >= (256ll << 20)) overlap_chunks_ = 8;
    if (effort_ < DTFFT_MEASURE) return DTFFT_SUCCESS;

    // timed choice on scratch buffers: forward + backward execute, max over ranks of the mean
    const size_t bytes = alloc_bytes();
    void *a = nullptr, *b = nullptr;
    int rc = mem_alloc(bytes, &a);
    if (!rc) rc = mem_alloc(bytes, &b);
    const bool saved_graphs = graphs_enabled_;
    graphs_enabled_ = false;
    int best = 1;
    double best_ms = 1e30;
    if (!rc) {
        cudaMemsetAsync(a, 0, bytes, stream_);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        for (int cand : {1, 4, 8}) {
            overlap_chunks_ = cand;
            auto pass = [&]() {
                int r2 = execute(a, b, DTFFT_EXECUTE_FORWARD, nullptr);
                if (r2) return r2;
                return execute(b, a, DTFFT_EXECUTE_BACKWARD, nullptr);
            };
            for (int i = 0; i < cfg_.n_measure_warmup_iters && !rc; ++i) rc = pass();
            cudaEventRecord(e0, stream_);
            for (int i = 0; i < cfg_.n_measure_iters && !rc; ++i) rc = pass();
            cudaEventRecord(e1, stream_);
            float t = 0;
            const bool ok = !rc && cudaEventSynchronize(e1) == cudaSuccess && cudaEventElapsedTime(&t, e0, e1) == cudaSuccess;
            const double ms = comm_.max(ok ? (double)t / std::max(1, (int)cfg_.n_measure_iters) : 1e30);
            log("autotune stage overlap, %d chunk(s): %.4f ms per forward + backward", cand, ms);
            if (ms < best_ms) best_ms = ms, best = cand;
            if (rc) break;
        }
        cudaEventDestroy(e0), cudaEventDestroy(e1);
    }
    graphs_enabled_ = saved_graphs;
    overlap_chunks_ = best;
    if (b) mem_free(b);
    if (a) mem_free(a);
    log("stage overlap: %d chunk(s)", overlap_chunks_);
    return rc;
}

int Plan::create_ffts() {
    // alloc_fft_plans + create_c2c_core / create_r2c_internal, dtfft_plan.F90:2221-2284, 2543-2556, 2636-2639
    const int me = comm_.rank(), nd = ndims_;
    for (int d = 0; d < nd; ++d) fft_mapping_[d] = d;
    if (!is_z_slab_ && !is_y_slab_) {
        for (int d = 0; d < nd; ++d)
            for (int d2 = 0; d2 < d; ++d2) {
                if (kind_ == PLAN_R2C && (d == 0 || d2 == 0)) continue;
                const Pencil &a = pencils_[(size_t)me][(size_t)d], &b = pencils_[(size_t)me][(size_t)d2];
                if (a.counts[0] == b.counts[0] && a.size() == b.size()) {
                    fft_mapping_[d] = fft_mapping_[d2];
                    break;
                }
            }
    }
    const int start = kind_ == PLAN_R2C ? 1 : 0;
    for (int d = start; d < nd; ++d) {
        int rank = 1;
        if ((is_z_slab_ && d == 0) || (is_y_slab_ && d == 1)) rank = 2;
        if ((is_z_slab_ && d == 1) || (is_y_slab_ && d == 2)) continue;
        const int m = fft_mapping_[d];
        if (fft_[m] && fft_[m]->created()) continue;
        fft_[m].reset(new FftExecutor);
        int rc = fft_[m]->create(rank, false, precision_, nullptr, pencils_[(size_t)me][(size_t)d], stream_);
        if (rc) return rc;
    }
    if (kind_ == PLAN_R2C) {
        fft_[0].reset(new FftExecutor);
        int rc = fft_[0]->create(is_z_slab_ ? 2 : 1, true, precision_, &real_pencils_[(size_t)me], pencils_[(size_t)me][0], stream_);
        if (rc) return rc;
    }
    return DTFFT_SUCCESS;
}

// ======================================================================================
// Plan: sizes
// ======================================================================================
size_t Plan::element_size() const { return (size_t)(kind_ == PLAN_R2C ? base_storage_ / 2 : base_storage_); }

int Plan::get_local_sizes(int32_t* in_starts, int32_t* in_counts, int32_t* out_starts, int32_t* out_counts,
                          size_t* alloc) const {
    // dtfft_plan.F90:1797-1877 + get_local_sizes, src/dtfft_pencil.F90:438-463
    const int me = comm_.rank(), nd = ndims_;
    const auto& pz = pencils_[(size_t)me];
    int out_dim = nd - 1;
    if (is_y_slab_ && nd == 3) out_dim = 1;
    auto vol = [&](const Pencil& p) { return (long long)p.size(); };
    long long internal = 0;
    for (int d = 0; d < nd; ++d) internal = std::max(internal, vol(pz[(size_t)d]));
    if (kind_ == PLAN_R2C) internal = std::max(vol(real_pencils_[(size_t)me]), 2 * internal);
    const Pencil& in_p = kind_ == PLAN_R2C ? real_pencils_[(size_t)me] : pz[0];
    const Pencil* ip = &in_p;
    const Pencil* op = &pz[(size_t)out_dim];
    long long total = internal;
    if (is_reshape_enabled_) {
        const Pencil& b1 = bricks_[(size_t)me][0];
        const Pencil& b2 = bricks_[(size_t)me][1];
        ip = &b1;
        long long a1 = vol(b1), a3 = vol(b2);
        if (kind_ == PLAN_R2C) a3 *= 2;
        total = std::max(total, std::max(a1, a3));
        if (is_final_reshape_enabled_) op = &b2;
    }
    for (int d = 0; d < nd; ++d) {
        if (in_starts) in_starts[d] = ip->starts[d];
        if (in_counts) in_counts[d] = ip->counts[d];
        if (out_starts) out_starts[d] = op->starts[d];
        if (out_counts) out_counts[d] = op->counts[d];
    }
    if (alloc) *alloc = (size_t)total;
    return DTFFT_SUCCESS;
}

size_t Plan::alloc_size() const {
    size_t a = 0;
    get_local_sizes(nullptr, nullptr, nullptr, nullptr, &a);
    return a;
}

size_t Plan::aux_bytes_transpose() const {
    size_t a = 0;
    if (dry_ && backend_is_pipelined(backend_)) {  // abstract_backend.F90:196-201
        for (int t : transpose_types()) {
            ExchangeDescription d;
            if (describe_exchange(t, &d) || d.members.size() < 2) continue;
            long long s = 0, r = 0;
            for (auto c : d.geo.send_counts) s += c;
            for (auto c : d.geo.recv_counts) r += c;
            a = std::max(a, (size_t)(std::max(s, r) * base_storage_));
        }
        return a;
    }
    for (auto& kv : handles_) a = std::max(a, (size_t)kv.second->aux_bytes());
    return a;
}

size_t Plan::aux_bytes_reshape() const {
    size_t a = 0;
    if (dry_ && is_reshape_enabled_ && (reshape_backend_ == BACKEND_NCCL || reshape_backend_ == BACKEND_NCCL_PIPELINED)) {
        // what the handles would report: pipelined workspace (abstract_backend.F90:196-201) and the
        // pack-free / unpack-free staging buffer (reshape_handle_generic.F90:684-686)
        const char* sc = getenv("DTFFTB_RESHAPE_SHORTCUTS");
        const bool shortcuts = !(sc && sc[0] == '0');
        for (int t = R_X_BRICKS_TO_PENCILS; t <= R_Z_BRICKS_TO_PENCILS; ++t) {
            HandleSpec hs;
            if (handle_spec(t, &hs) || hs.members.size() < 2) continue;
            const ReshapeGeometry g = reshape_geometry(t, hs.send, hs.recv, hs.me);
            long long s = 0, r = 0;
            for (auto c : g.send_counts) s += c;
            for (auto c : g.recv_counts) r += c;
            if (backend_is_pipelined(reshape_backend_)) a = std::max(a, (size_t)(std::max(s, r) * hs.es));
            if (shortcuts && (g.is_pack_free || g.is_unpack_free))
                a = std::max(a, (size_t)(std::max(hs.send[(size_t)hs.me].size(), hs.recv[(size_t)hs.me].size()) * hs.es));
        }
        return a;
    }
    for (auto& kv : rhandles_) a = std::max(a, (size_t)kv.second->aux_bytes());
    return a;
}

size_t Plan::aux_bytes() const {  // dtfft_plan.F90:1398-1420
    return std::max(aux_bytes_transpose(), aux_bytes_reshape()) + alloc_bytes();
}

int Plan::get_pencil(int layout, dtfft_pencil_t* p) const {
    // dtfft_plan.F90:1286-1336
    if (layout < DTFFT_LAYOUT_X_BRICKS || layout > DTFFT_LAYOUT_Z_BRICKS) return DTFFT_ERROR_INVALID_LAYOUT;
    const int me = comm_.rank();
    bool valid = true;
    if (!is_reshape_enabled_ && (layout == DTFFT_LAYOUT_X_BRICKS || layout == DTFFT_LAYOUT_Z_BRICKS)) valid = false;
    if (ndims_ == 2 && layout == DTFFT_LAYOUT_Z_PENCILS) valid = false;
    if (layout == DTFFT_LAYOUT_X_PENCILS_FOURIER && kind_ != PLAN_R2C) valid = false;
    if (!valid) return DTFFT_ERROR_INVALID_LAYOUT;
    const Pencil* src = nullptr;
    switch (layout) {
        case DTFFT_LAYOUT_X_BRICKS: src = &bricks_[(size_t)me][0]; break;
        case DTFFT_LAYOUT_X_PENCILS: src = kind_ == PLAN_R2C ? &real_pencils_[(size_t)me] : &pencils_[(size_t)me][0]; break;
        case DTFFT_LAYOUT_X_PENCILS_FOURIER: src = &pencils_[(size_t)me][0]; break;
        case DTFFT_LAYOUT_Y_PENCILS: src = &pencils_[(size_t)me][1]; break;
        case DTFFT_LAYOUT_Z_PENCILS: src = &pencils_[(size_t)me][2]; break;
        default: src = &bricks_[(size_t)me][1]; break;
    }
    std::memset(p, 0, sizeof(*p));
    p->dim = (uint8_t)src->aligned_dim;
    p->ndims = (uint8_t)ndims_;
    for (int d = 0; d < ndims_; ++d) p->starts[d] = src->starts[d], p->counts[d] = src->counts[d];
    p->size = (size_t)src->size();
    return DTFFT_SUCCESS;
}

// ======================================================================================
// Plan: memory (alloc_mem / free_mem, src/dtfft_reshape_plan_base.F90:393-532)
// ======================================================================================
int Plan::mem_alloc(size_t bytes, void** ptr) {
    if (!ptr) return DTFFT_ERROR_INVALID_USAGE;
    *ptr = nullptr;
    if (dry_) return DTFFT_ERROR_GPU_NOT_SET;
    if (bytes == 0) return DTFFT_ERROR_INVALID_ALLOC_BYTES;
    const bool fused = backend_ == BACKEND_NVLINK_FUSED || reshape_backend_ == BACKEND_NVLINK_FUSED;
    // With the fused backend the registration below is collective (an allgather of IPC handles): a rank whose
    // allocation fails must not leave the others waiting in it.  The local outcome is agreed on first and
    // every rank backs out together.
    const bool collective = fused && comm_.size() > 1 && peers_.available();
    int local_rc = DTFFT_SUCCESS;
    {  // reshape_plan_base.F90:412, 429-432: refuse what cannot fit instead of letting the allocator fail
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            if (bytes > free_b) local_rc = DTFFT_ERROR_ALLOC_FAILED;
        } else {
            cudaGetLastError();
        }
    }
    if (local_rc && !collective) return local_rc;
    Alloc a{};
    a.bytes = bytes;
    const bool use_nccl_alloc = nccl_ && !fused && (backend_ == BACKEND_NCCL || backend_ == BACKEND_NCCL_PIPELINED) &&
                                getenv("DTFFTB_NO_NCCL_MEM") == nullptr;
    if (use_nccl_alloc && !local_rc) {
        if (ncclMemAlloc(&a.ptr, bytes) == ncclSuccess) {
            a.nccl = true;
            if (ncclCommRegister(nccl_, a.ptr, bytes, &a.reg) != ncclSuccess) a.reg = nullptr;
        } else {
            a.ptr = nullptr;
        }
    }
    if (!a.ptr && !local_rc) {
        cudaError_t ce = cudaMalloc(&a.ptr, bytes);
        if (ce != cudaSuccess) {
            cudaGetLastError();
            a.ptr = nullptr;
            local_rc = DTFFT_ERROR_ALLOC_FAILED;
        }
    }
    if (collective) {
        if (comm_.sum(local_rc == DTFFT_SUCCESS ? 1 : 0) != comm_.size()) {
            if (a.ptr) cudaFree(a.ptr);
            cudaGetLastError();
            return DTFFT_ERROR_ALLOC_FAILED;
        }
        int slot = -1;
        int rc = peers_.register_buffer(a.ptr, bytes, &slot);
        if (rc) return rc;
        a.peer = slot >= 0;
    } else if (local_rc) {
        return local_rc;
    }
    allocs_.push_back(a);
    *ptr = a.ptr;
    return DTFFT_SUCCESS;
}

int Plan::mem_free(void* ptr) {
    for (size_t i = 0; i < allocs_.size(); ++i) {
        if (allocs_[i].ptr != ptr) continue;
        Alloc a = allocs_[i];
        allocs_.erase(allocs_.begin() + (long)i);
        forget_buffer_caches();
        if (a.peer) peers_.unregister_buffer(a.ptr);
        if (a.nccl) {
            if (a.reg) ncclCommDeregister(nccl_, a.reg);
            if (ncclMemFree(a.ptr) != ncclSuccess) return DTFFT_ERROR_FREE_FAILED;
        } else if (cudaFree(a.ptr) != cudaSuccess) {
            cudaGetLastError();
            return DTFFT_ERROR_FREE_FAILED;
        }
        // the memory is gone either way; tell the caller if the plan died of a peer time-out
        return peers_.error_state() ? DTFFTB_ERROR_PEER_TIMEOUT : DTFFT_SUCCESS;
    }
    return DTFFT_ERROR_FREE_FAILED;
}

int Plan::register_buffer(void* ptr, size_t bytes) {
    if (!peers_.available() || comm_.size() == 1) return DTFFT_SUCCESS;
    int slot = -1;
    int rc = peers_.register_buffer(ptr, bytes, &slot);
    if (rc) return rc;
    return slot >= 0 ? DTFFT_SUCCESS : DTFFTB_ERROR_NOT_REGISTERED;
}

int Plan::unregister_buffer(void* ptr) {
    forget_buffer_caches();
    return peers_.unregister_buffer(ptr);
}

// Everything keyed by a buffer ADDRESS (captured graphs, fused kernels holding peer mappings of
// that address) dies with the buffer: a later allocation may reuse the address with other peers.
void Plan::forget_buffer_caches() {
    if (stream_) cudaStreamSynchronize(stream_);
    drop_graphs();
    for (auto& kv : handles_) kv.second->forget_buffers();
    for (auto& kv : rhandles_) kv.second->forget_buffers();
}

int Plan::check_aux(void* aux, bool from_execute, void** aux1, void** aux2) {
    // dtfft_plan.F90:2286-2331
    const size_t shift = alloc_bytes();
    const bool need2 = aux_bytes_transpose() > 0 || aux_bytes_reshape() > 0;
    *aux2 = nullptr;
    if (!is_aux_alloc_ && !aux) {
        int rc = mem_alloc(aux_bytes(), &aux_ptr_);
        if (rc) return rc;
        is_aux_alloc_ = true;
    }
    *aux1 = is_aux_alloc_ ? aux_ptr_ : aux;
    if (from_execute && need2) *aux2 = static_cast<char*>(*aux1) + shift;
    return DTFFT_SUCCESS;
}

int Plan::check_device_ptrs(const void* a, const void* b, const void* c) const {
    // check_device_pointers, dtfft_plan.F90:1769-1795 (is_device_ptr, src/dtfft_helpers.c:9-16)
    const void* ps[3] = {a, b, c};
    for (const void* p : ps) {
        if (!p) continue;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
            cudaGetLastError();
            return DTFFT_ERROR_NOT_DEVICE_PTR;
        }
        if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) return DTFFT_ERROR_NOT_DEVICE_PTR;
    }
    return DTFFT_SUCCESS;
}

// ======================================================================================
// Plan: execution
// ======================================================================================
int Plan::run_transpose(int ttype, void* in, void* out, void* aux) {
    auto it = handles_.find(ttype);
    if (it == handles_.end()) return DTFFT_ERROR_INVALID_TRANSPOSE_TYPE;
    ReshapeHandle& h = *it->second;
    static const char* const names[7] = {"Transpose Z_TO_X", "Transpose Z_TO_Y", "Transpose Y_TO_X", "", "Transpose X_TO_Y",
                                         "Transpose Y_TO_Z", "Transpose X_TO_Z"};  // reshape_plan_base.F90:229
    TraceRange trace(names[ttype + 3], kColorTransposeType[ttype + 3]);
    stat_launches_ += h.kernel_launches();
    stat_local_ += h.local_elements() * base_storage_;
    stat_remote_ += h.remote_elements() * base_storage_;
    const int rc = h.execute(in, out, stream_, aux);
    if (rc == DTFFTB_ERROR_NOT_REGISTERED && h.backend() == BACKEND_NVLINK_FUSED) return fallback_execute(false, ttype, in, out, aux);
    return rc;
}

// A destination that cudaIpc cannot share (agreed on by every rank inside publish): this call -- and every
// later one with such a buffer -- runs on the NCCL backend instead.  Logged once; CUDA-graph replay is
// switched off for the plan because NCCL calls are kept out of graphs.
int Plan::fallback_execute(bool reshape, int type, void* in, void* out, void* aux) {
    auto& fb = reshape ? fb_rhandles_ : fb_handles_;
    if (fb.empty()) {
        if (!nccl_) return DTFFTB_ERROR_NOT_REGISTERED;
        HandleContext ctx;
        ctx.nccl = nccl_;
        ctx.peers = &peers_;
        ctx.effort = effort_;
        ctx.no_shortcuts = true;  // the three-step schedule needs no workspace
        std::vector<int> types;
        if (reshape)
            for (int t = R_X_BRICKS_TO_PENCILS; t <= R_Z_BRICKS_TO_PENCILS; ++t) types.push_back(t);
        else
            types = transpose_types();
        for (int t : types) {
            HandleSpec hs;
            int rc = handle_spec(t, &hs);
            if (rc) return rc;
            std::unique_ptr<ReshapeHandle> h(new ReshapeHandle);
            const int b = hs.members.size() > 1 ? BACKEND_NCCL : BACKEND_NONE;
            rc = h->create(ctx, reshape ? 0 : t, reshape ? t : 0, hs.comm_id, hs.members, hs.me, hs.send, hs.recv, hs.es, b);
            if (rc) return rc;
            fb[t] = std::move(h);
        }
        log("NVLINK_FUSED: a buffer cannot be shared through cudaIpc; such calls use the NCCL backend");
        if (graphs_enabled_) {
            graphs_enabled_ = false;
            if (!graphs_.empty()) {
                cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
                cudaStreamIsCapturing(stream_, &st);
                if (st == cudaStreamCaptureStatusNone) drop_graphs();
            }
        }
    }
    auto it = fb.find(type);
    if (it == fb.end()) return DTFFTB_ERROR_INTERNAL;
    ++stat_fallbacks_;
    return it->second->execute(in, out, stream_, aux);
}

int Plan::run_reshape(int rtype, void* in, void* out, void* aux) {
    auto it = rhandles_.find(rtype);
    if (it == rhandles_.end()) return DTFFT_ERROR_INVALID_RESHAPE_TYPE;
    ReshapeHandle& h = *it->second;
    const int64_t es = (rtype == R_X_BRICKS_TO_PENCILS || rtype == R_X_PENCILS_TO_BRICKS) ? base_storage_init_ : base_storage_;
    static const char* const names[4] = {"Reshape X_BRICKS_TO_PENCILS", "Reshape X_PENCILS_TO_BRICKS",
                                         "Reshape Z_PENCILS_TO_BRICKS", "Reshape Z_BRICKS_TO_PENCILS"};
    TraceRange trace(names[rtype - R_X_BRICKS_TO_PENCILS], kColorReshapeType[rtype - R_X_BRICKS_TO_PENCILS]);
    stat_launches_ += h.kernel_launches();
    stat_local_ += h.local_elements() * es;
    stat_remote_ += h.remote_elements() * es;
    const int rc = h.execute(in, out, stream_, aux);
    if (rc == DTFFTB_ERROR_NOT_REGISTERED && h.backend() == BACKEND_NVLINK_FUSED) return fallback_execute(true, rtype, in, out, aux);
    return rc;
}

int Plan::run_fft(int dim, void* a, void* b, int sign) {
    FftExecutor* f = fft_[fft_mapping_[dim]].get();
    if (!f) return DTFFT_SUCCESS;
    TraceRange trace("FFT", kColorFft);  // abstract_executor.F90:230
    return f->execute(a, b, sign);
}

int Plan::run_fft_transpose(int dim, void* a, void* b, int sign, int ttype, void* c, void* aux) {
    FftExecutor* f = fft_[fft_mapping_[dim]].get();
    auto it = handles_.find(ttype);
    if (it == handles_.end()) return DTFFT_ERROR_INVALID_TRANSPOSE_TYPE;
    ReshapeHandle& h = *it->second;
    // Peers store into my `c` while I still transform later chunks: `c` must not be the FFT's
    // source (`b` never is: a transposition is out of place).  Every rank of the group takes
    // the same decision except ranks without data, whose barriers still pair up.
    long long nch = overlap_chunks_;
    const long long slow = h.slow_extent();
    bool ok = nch > 1 && f && f->created() && h.can_chunk() && c != a && slow > 1 && f->how_many() % slow == 0;
    if (ok) nch = std::min(nch, slow);
    if (!ok || nch <= 1) {
        int rc = run_fft(dim, a, b, sign);
        if (rc) return rc;
        return run_transpose(ttype, b, c, aux);
    }
    cudaError_t ce;
    if (!xfer_stream_) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = greatest priority
        ce = cudaStreamCreateWithPriority(&xfer_stream_, cudaStreamNonBlocking, hi);
        if (ce != cudaSuccess) return cuda_error(ce);
        ce = cudaEventCreateWithFlags(&xfer_done_, cudaEventDisableTiming);
        if (ce != cudaSuccess) return cuda_error(ce);
    }
    while ((long long)chunk_events_.size() < nch) {
        cudaEvent_t e;
        ce = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        if (ce != cudaSuccess) return cuda_error(ce);
        chunk_events_.push_back(e);
    }
    const long long per_slow = f->how_many() / slow;
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int ctas = overlap_ctas_ > 0 ? overlap_ctas_ : sms;
    int rc = h.fused_begin(c, stream_);
    if (rc == DTFFTB_ERROR_NOT_REGISTERED) {  // unshareable destination: sequential, NCCL stand-in
        rc = run_fft(dim, a, b, sign);
        if (rc) return rc;
        return run_transpose(ttype, b, c, aux);
    }
    if (rc) return rc;
    for (long long k = 0; k < nch; ++k) {
        const long long lo = slow * k / nch, hi = slow * (k + 1) / nch;
        rc = f->execute_range(a, b, sign, lo * per_slow, (hi - lo) * per_slow);
        if (rc) return rc;
        ce = cudaEventRecord(chunk_events_[(size_t)k], stream_);
        if (ce != cudaSuccess) return cuda_error(ce);
        ce = cudaStreamWaitEvent(xfer_stream_, chunk_events_[(size_t)k], 0);
        if (ce != cudaSuccess) return cuda_error(ce);
        // the last chunk has nothing to hide behind: let it use the whole GPU
        rc = h.fused_chunk(b, c, (int)k, (int)nch, k + 1 < nch ? ctas : 0, xfer_stream_);
        if (rc) return rc;
    }
    ce = cudaEventRecord(xfer_done_, xfer_stream_);
    if (ce != cudaSuccess) return cuda_error(ce);
    ce = cudaStreamWaitEvent(stream_, xfer_done_, 0);
    if (ce != cudaSuccess) return cuda_error(ce);
    rc = h.fused_end(stream_);
    if (rc) return rc;
    stat_launches_ += 2 + nch;
    stat_local_ += h.local_elements() * base_storage_;
    stat_remote_ += h.remote_elements() * base_storage_;
    stat_overlapped_ += 1, stat_eager_only_ += 1;
    return DTFFT_SUCCESS;
}

// Two consecutive transpositions of a transpose-only schedule, a -> b -> c.  On slab-shaped process grids one of
// them is local (one rank in its communicator, HBM-bound) and the other exchanges over NVLink; when the exchange
// runs in its copy-engine form the two are pipelined peer by peer (below), otherwise they run one after the other.
// (A chunk-wise pipeline of the direct-store kernel was measured too and dropped: the remote stores of a kernel
// collapse to ~420 GB/s when an HBM-bound kernel runs beside them, profiles/r02c_nvlink_probe_n2.md, so it gained
// 1.4 % at best, profiles/r02b_bench_n2_pair*.json.)  No reference counterpart.
int Plan::run_transpose_pair(int t1, void* a, void* b, int t2, void* c, void* aux) {
    auto i1 = handles_.find(t1), i2 = handles_.find(t2);
    if (i1 == handles_.end() || i2 == handles_.end()) return DTFFT_ERROR_INVALID_TRANSPOSE_TYPE;
    ReshapeHandle &h1 = *i1->second, &h2 = *i2->second;
    const bool distinct = a != b && b != c && a != c;
    // Copy-engine exchanges pipeline PEER BY PEER with the local transposition next to them (the copies keep their
    // rate while SM kernels use the HBM, profiles/r02c_nvlink_probe_n2.md; the direct-store kernel does not):
    //   producer: [local piece for peer p -> pack p] on the plan stream, copy p on a copy stream, for every peer;
    //   consumer: as each sender's block lands (pairwise flag behind its copy) the local piece that reads it runs.
    // (every member must take the same path: nobody may be without data)
    if (pair_overlap_ && distinct && aux && h1.is_local_transpose() && h2.dma_mode() && h2.min_member_slow_extent() > 0) {
        TraceRange trace("Transpose pair (local pieces -> packs || copies)", kColorTranspose);
        const int P = h2.n_members(), me = h2.my_index();
        int rc = h2.dma_begin(c, stream_);
        if (rc == DTFFTB_ERROR_NOT_REGISTERED) {
            rc = run_transpose(t1, a, b, aux);
            if (rc) return rc;
            return run_transpose(t2, b, c, aux);
        }
        if (rc) return rc;
        const Pencil& mine = h2.send_by_member()[(size_t)me];  // my pencil between the two transpositions
        for (int k = 1; k < P; ++k) {
            const int p = (me + k) % P;
            const int nsub = h2.n_subs_to(p);
            for (int q = 0; q < nsub; ++q) {
                rc = h1.local_piece(a, b, 0, mine, h2.recv_by_member()[(size_t)p], p, q, nsub, stream_);
                if (rc) return rc;
                rc = h2.dma_send(b, c, aux, p, q, stream_);
                if (rc) return rc;
            }
        }
        rc = h1.local_piece(a, b, 0, mine, h2.recv_by_member()[(size_t)me], me, 0, 1, stream_);
        if (rc) return rc;
        rc = h2.dma_self(b, c, stream_);
        if (rc) return rc;
        rc = h2.dma_end(stream_, true);
        if (rc) return rc;
        stat_launches_ += 2 + 3 * P;
        stat_local_ += (h1.local_elements() + h2.local_elements()) * base_storage_;
        stat_remote_ += (h1.remote_elements() + h2.remote_elements()) * base_storage_;
        // replayed from a CUDA graph this pipeline loses part of its overlap (2 B200: 2.06 vs 1.91 ms per cycle,
        // profiles/r02f_bench_n2_dma*.json), like the FFT stage overlap did: it stays eager
        stat_overlapped_ += 1, stat_eager_only_ += 1;
        return DTFFT_SUCCESS;
    }
    if (pair_overlap_ && distinct && aux && h1.dma_mode() && h2.is_local_transpose() && h1.min_member_slow_extent() > 0) {
        TraceRange trace("Transpose pair (packs || copies -> local pieces)", kColorTranspose);
        const int P = h1.n_members(), me = h1.my_index();
        int rc = h1.dma_begin(b, stream_);
        if (rc == DTFFTB_ERROR_NOT_REGISTERED) {
            rc = run_transpose(t1, a, b, aux);
            if (rc) return rc;
            return run_transpose(t2, b, c, aux);
        }
        if (rc) return rc;
        rc = h1.dma_advance_signals(stream_);  // new epoch of the pairwise "landed" flags, before any copy is enqueued
        if (rc) return rc;
        const Pencil& mine = h1.recv_by_member()[(size_t)me];  // my pencil between the two transpositions
        for (int k = 1; k < P; ++k) {
            const int p = (me + k) % P;
            for (int q = 0; q < h1.n_subs_to(p); ++q) {
                rc = h1.dma_send(a, b, aux, p, q, stream_);
                if (rc) return rc;
                rc = h1.dma_signal(p, q);
                if (rc) return rc;
            }
        }
        rc = h1.dma_self(a, b, stream_);
        if (rc) return rc;
        rc = h2.local_piece(b, c, 1, h1.send_by_member()[(size_t)me], mine, me, 0, 1, stream_);
        if (rc) return rc;
        for (int k = 1; k < P; ++k) {  // sender (me - k) has me as its k-th target: blocks arrive in this order
            const int r = (me - k + P) % P;
            const int nsub = h1.n_subs_from(r);
            for (int q = 0; q < nsub; ++q) {
                rc = h1.dma_wait(r, q, stream_);
                if (rc) return rc;
                rc = h2.local_piece(b, c, 1, h1.send_by_member()[(size_t)r], mine, r, q, nsub, stream_);
                if (rc) return rc;
            }
        }
        rc = h1.dma_end(stream_, false);  // my copies are done with the staging buffer; no group barrier needed
        if (rc) return rc;
        stat_launches_ += 1 + 5 * P;
        stat_local_ += (h1.local_elements() + h2.local_elements()) * base_storage_;
        stat_remote_ += (h1.remote_elements() + h2.remote_elements()) * base_storage_;
        // replayed from a CUDA graph this pipeline loses part of its overlap (2 B200: 2.06 vs 1.91 ms per cycle,
        // profiles/r02f_bench_n2_dma*.json), like the FFT stage overlap did: it stays eager
        stat_overlapped_ += 1, stat_eager_only_ += 1;
        return DTFFT_SUCCESS;
    }
    int rc = run_transpose(t1, a, b, aux);
    if (rc) return rc;
    return run_transpose(t2, b, c, aux);
}

// Introspection for host tests: the copy-engine form of one transposition on this rank (geometry.h: DmaBlock)
// and the pieces of a local transposition cut by the members of the exchange next to it.
int Plan::describe_dma(int ttype, std::vector<int>* members, int* me, std::vector<DmaEntry>* entries) const {
    HandleSpec hs;
    int rc = handle_spec(ttype, &hs);
    if (rc) return rc;
    if (members) *members = hs.members;
    if (me) *me = hs.me;
    entries->clear();
    const Pencil& send = hs.send[(size_t)hs.me];
    long long off = 0;
    for (size_t i = 0; i < hs.members.size(); ++i) {
        const int nsub = (int)i == hs.me ? 1 : dma_nsub(send, hs.recv[i], hs.es, (int)hs.members.size());
        for (int q = 0; q < nsub; ++q) {
            DmaEntry e;
            e.member = (int)i, e.sub = q, e.nsub = nsub;
            bool tr = false;
            e.fused = block_box(send, hs.recv[i], q, nsub, &tr);
            if ((int)i == hs.me || e.fused.empty()) {
                e.blk.ok = true;
            } else {
                e.blk = dma_block(e.fused, tr, off);
                off += e.fused.volume();
            }
            entries->push_back(e);
        }
    }
    return DTFFT_SUCCESS;
}

int Plan::describe_peer_piece(int t_local, int t_exchange, int side, int peer, int sub, int* nsub_out, Box* box) const {
    if (side != 0 && side != 1) return DTFFT_ERROR_INVALID_USAGE;
    HandleSpec hl, hx;
    int rc = handle_spec(t_local, &hl);
    if (rc) return rc;
    rc = handle_spec(t_exchange, &hx);
    if (rc) return rc;
    if (hl.members.size() != 1) return DTFFT_ERROR_INVALID_USAGE;  // not a local transposition
    if (peer < 0 || peer >= (int)hx.members.size()) return DTFFT_ERROR_INVALID_USAGE;
    // side 0: the exchange follows: block (my pencil in between -> the peer's destination);
    // side 1: it precedes: block (the sender's source -> my pencil in between)
    const Pencil& x_src = side == 0 ? hx.send[(size_t)hx.me] : hx.send[(size_t)peer];
    const Pencil& x_dst = side == 0 ? hx.recv[(size_t)peer] : hx.recv[(size_t)hx.me];
    const int nsub = peer == hx.me ? 1 : dma_nsub(x_src, x_dst, hx.es, (int)hx.members.size());
    if (nsub_out) *nsub_out = nsub;
    if (sub < 0 || sub >= nsub) return DTFFT_ERROR_INVALID_USAGE;
    *box = local_box_for_block(hl.send[0], hl.recv[0], x_src, x_dst, sub, nsub);
    return DTFFT_SUCCESS;
}

int Plan::transpose(void* in, void* out, int ttype, void* aux) {
    // transpose_private, dtfft_plan.F90:695-747
    TraceRange trace_api("dtfft_transpose", kColorTranspose);
    if (!created_) return DTFFT_ERROR_PLAN_NOT_CREATED;
    if (dry_) return DTFFT_ERROR_GPU_NOT_SET;
    if (peers_.error_state()) return DTFFTB_ERROR_PEER_TIMEOUT;  // a device barrier timed out earlier: results are void
    const int at = std::abs(ttype);
    if (at < 1 || at > 3 || (ndims_ == 2 && at > 1) || (at == 3 && !is_z_slab_)) return DTFFT_ERROR_INVALID_TRANSPOSE_TYPE;
    if (in == out) return DTFFT_ERROR_INPLACE_TRANSPOSE;
    if (in == aux || out == aux) return DTFFT_ERROR_INVALID_AUX;
    int rc = check_device_ptrs(in, out, aux);
    if (rc) return rc;
    stat_launches_ = stat_local_ = stat_remote_ = 0;
    void *a1 = nullptr, *a2 = nullptr;
    if (aux_bytes_transpose() > 0 || aux) {
        rc = check_aux(aux, false, &a1, &a2);
        if (rc) return rc;
    }
    return run_transpose(ttype, in, out, a1);
}

int Plan::reshape(void* in, void* out, int rtype, void* aux) {
    // reshape_private, dtfft_plan.F90:489-544
    TraceRange trace_api("dtfft_reshape", kColorTranspose);
    if (!created_) return DTFFT_ERROR_PLAN_NOT_CREATED;
    if (dry_) return DTFFT_ERROR_GPU_NOT_SET;
    if (peers_.error_state()) return DTFFTB_ERROR_PEER_TIMEOUT;  // a device barrier timed out earlier: results are void
    if (!is_reshape_enabled_) return DTFFT_ERROR_RESHAPE_NOT_SUPPORTED;
    if (rtype < R_X_BRICKS_TO_PENCILS || rtype > R_Z_BRICKS_TO_PENCILS) return DTFFT_ERROR_INVALID_RESHAPE_TYPE;
    if (in == out) return DTFFT_ERROR_INPLACE_RESHAPE;
    if (in == aux || out == aux) return DTFFT_ERROR_INVALID_AUX;
    int rc = check_device_ptrs(in, out, aux);
    if (rc) return rc;
    stat_launches_ = stat_local_ = stat_remote_ = 0;
    void *a1 = nullptr, *a2 = nullptr;
    if (aux_bytes_reshape() > 0 || aux) {
        rc = check_aux(aux, false, &a1, &a2);
        if (rc) return rc;
    }
    return run_reshape(rtype, in, out, a1);
}

int Plan::execute(void* in, void* out, int execute_type, void* aux) {
    // execute_ptr, dtfft_plan.F90:771-837
    TraceRange trace_api("dtfft_execute", kColorExecute);
    if (!created_) return DTFFT_ERROR_PLAN_NOT_CREATED;
    if (dry_) return DTFFT_ERROR_GPU_NOT_SET;
    if (peers_.error_state()) return DTFFTB_ERROR_PEER_TIMEOUT;  // a device barrier timed out earlier: results are void
    if (execute_type != DTFFT_EXECUTE_FORWARD && execute_type != DTFFT_EXECUTE_BACKWARD) return DTFFT_ERROR_INVALID_EXECUTE_TYPE;
    const bool inplace = in == out;
    if (is_transpose_plan_ && inplace &&
        (ndims_ == 2 || is_y_slab_ || (is_reshape_enabled_ && !is_final_reshape_enabled_)))
        return DTFFT_ERROR_INPLACE_TRANSPOSE;
    if (in == aux || out == aux) return DTFFT_ERROR_INVALID_AUX;
    if (is_transpose_plan_ && kind_ == PLAN_R2C) return DTFFT_ERROR_R2C_EXECUTE_CALLED;
    int rc = check_device_ptrs(in, out, aux);
    if (rc) return rc;
    stat_launches_ = stat_local_ = stat_remote_ = stat_overlapped_ = stat_eager_only_ = 0;
    void *a1 = nullptr, *a2 = nullptr;
    rc = check_aux(aux, true, &a1, &a2);
    if (rc) return rc;
    const bool fwd = execute_type == DTFFT_EXECUTE_FORWARD;
    if (!graphs_usable()) return execute_schedule(in, out, fwd, a1, a2, inplace);

    // CUDA-graph replay of the whole schedule (no reference counterpart; the reference enqueues
    // every kernel of every execute from the host).  First call with a given (in, out, aux,
    // direction): eager, so that every lazily built table / cuFFT plan / barrier group exists.
    // Second call: captured while it is enqueued.  Later calls: one cudaGraphLaunch.
    GraphKey key{in, out, a1, fwd};
    auto it = graphs_.find(key);
    // peers hold mappings of my buffers (NVLINK_FUSED): an address that was freed and re-allocated since the
    // graph was captured must go through the eager path again, which re-publishes it (handle.h: peer_bases)
    unsigned long long ids[3] = {0, 0, 0};
    if (comm_.size() > 1) {
        ids[0] = buffer_id(in), ids[1] = buffer_id(out), ids[2] = buffer_id(a1);
        if (it != graphs_.end() && (it->second.ids[0] != ids[0] || it->second.ids[1] != ids[1] || it->second.ids[2] != ids[2])) {
            if (it->second.exec) cudaGraphExecDestroy(it->second.exec);
            graphs_.erase(it);
            it = graphs_.end();
        }
    }
    if (it == graphs_.end()) {
        if (graphs_.size() >= 16) drop_graphs();
        it = graphs_.emplace(key, GraphEntry{}).first;
        for (int i = 0; i < 3; ++i) it->second.ids[i] = ids[i];
        rc = execute_schedule(in, out, fwd, a1, a2, inplace);
        if (!graphs_enabled_) return rc;  // the call fell back to NCCL and dropped the graphs: `it` is gone
        // A schedule with overlapped stages stays eager: measured on 2 B200 (profiles/r01f_configs_auto_n2.jsonl
        // vs r01d_configs_n2.jsonl, 16384^2 slab) the two-stream pipeline loses its overlap when replayed as
        // graph branches (9.41 ms vs 8.55 ms eager), and such schedules are long enough not to be launch-bound.
        if (stat_eager_only_ > 0) it->second.failed = true;
        return rc;
    }
    GraphEntry& g = it->second;
    if (g.failed) return execute_schedule(in, out, fwd, a1, a2, inplace);
    if (g.exec && g.evictions != handle_evictions()) {  // kernels the graph points at were destroyed
        cudaGraphExecDestroy(g.exec);
        g.exec = nullptr;
        return execute_schedule(in, out, fwd, a1, a2, inplace);  // eager: rebuilds them; next call re-captures
    }
    if (!g.exec) {
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(stream_, cudaStreamCaptureModeRelaxed);
        if (ce != cudaSuccess) {
            cudaGetLastError();
            g.failed = true;
            return execute_schedule(in, out, fwd, a1, a2, inplace);
        }
        rc = execute_schedule(in, out, fwd, a1, a2, inplace);
        ce = cudaStreamEndCapture(stream_, &graph);
        if (rc == DTFFT_SUCCESS && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&g.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc != DTFFT_SUCCESS || ce != cudaSuccess || !g.exec) {
            // nothing was enqueued by the failed capture: run this call eagerly and stop trying
            cudaGetLastError();
            g.exec = nullptr;
            g.failed = true;
            stat_launches_ = stat_local_ = stat_remote_ = stat_overlapped_ = stat_eager_only_ = 0;
            return execute_schedule(in, out, fwd, a1, a2, inplace);
        }
        g.launches = stat_launches_, g.local = stat_local_, g.remote = stat_remote_, g.overlapped = stat_overlapped_;
        g.evictions = handle_evictions();
    }
    stat_launches_ = g.launches, stat_local_ = g.local, stat_remote_ = g.remote, stat_overlapped_ = g.overlapped;
    cudaError_t ce = cudaGraphLaunch(g.exec, stream_);
    if (ce != cudaSuccess) return cuda_error(ce);
    ++stat_graph_replays_;
    return DTFFT_SUCCESS;
}

long long Plan::handle_evictions() const {
    long long n = 0;
    for (auto& kv : handles_) n += kv.second->evictions();
    for (auto& kv : rhandles_) n += kv.second->evictions();
    return n;
}

bool Plan::graphs_usable() const {
    if (!graphs_enabled_) return false;
    // NCCL collectives stay out of graphs: capturing them (measured on 2 B200, profiles/r02b_half_ncclgraphs*_n2.jsonl)
    // changed the launch-bound half-size configs by -1.7 % ... +3 %, not a win worth the user-buffer / proxy-thread caveats
    auto ok = [](int b) { return b == BACKEND_NONE || b == BACKEND_NVLINK_FUSED; };
    if (comm_.size() > 1 && !ok(backend_)) return false;
    if (comm_.size() > 1 && is_reshape_enabled_ && !ok(reshape_backend_)) return false;
    return true;
}

void Plan::drop_graphs() {
    for (auto& kv : graphs_)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    graphs_.clear();
}

int Plan::execute_schedule(void* in, void* out, bool fwd, void* a1, void* a2, bool inplace) {
    // execute_private, :839-875
    if (ndims_ == 2 || is_y_slab_)
        return is_reshape_enabled_ ? execute_2d_reshape(in, out, fwd, a1, a2) : execute_2d(in, out, fwd, a1, a2);
    if (is_z_slab_)
        return is_reshape_enabled_ ? execute_z_slab_reshape(in, out, fwd, a1, a2) : execute_z_slab(in, out, fwd, a1, inplace, a2);
    return is_reshape_enabled_ ? execute_generic_reshape(in, out, fwd, a1, a2) : execute_generic(in, out, fwd, a1, a2);
}

#define RUN(x)            \
    do {                  \
        int rc_ = (x);    \
        if (rc_) return rc_; \
    } while (0)

int Plan::execute_2d(void* in, void* out, bool fwd, void* aux, void* aux2) {  // :877-913
    const int last = 1;  // fft(fft_mapping(2))
    if (is_transpose_plan_) return run_transpose(fwd ? T_X_TO_Y : T_Y_TO_X, in, out, aux);
    if (fwd) {
        RUN(run_fft_transpose(0, in, aux, -1, T_X_TO_Y, out, aux2));
        RUN(run_fft(last, out, out, -1));
    } else {
        RUN(run_fft_transpose(last, in, in, +1, T_Y_TO_X, aux, aux2));
        RUN(run_fft(0, aux, out, +1));
    }
    return DTFFT_SUCCESS;
}

int Plan::execute_2d_reshape(void* in, void* out, bool fwd, void* aux, void* aux2) {  // :915-972
    if (is_transpose_plan_) {
        if (fwd) {
            RUN(run_reshape(R_X_BRICKS_TO_PENCILS, in, aux, aux2));
            if (is_final_reshape_enabled_) {
                RUN(run_transpose(T_X_TO_Y, aux, in, aux2));
                RUN(run_reshape(R_Z_PENCILS_TO_BRICKS, in, out, aux2));
            } else {
                RUN(run_transpose(T_X_TO_Y, aux, out, aux2));
            }
        } else {
            if (is_final_reshape_enabled_) {
                RUN(run_reshape(R_Z_BRICKS_TO_PENCILS, in, out, aux2));
                RUN(run_transpose(T_Y_TO_X, out, aux, aux2));
            } else {
                RUN(run_transpose(T_Y_TO_X, in, aux, aux2));
            }
            RUN(run_reshape(R_X_PENCILS_TO_BRICKS, aux, out, aux2));
        }
        return DTFFT_SUCCESS;
    }
    if (fwd) {
        RUN(run_reshape(R_X_BRICKS_TO_PENCILS, in, aux, aux2));
        RUN(run_fft_transpose(0, aux, out, -1, T_X_TO_Y, aux, aux2));
        if (is_final_reshape_enabled_) {
            RUN(run_fft(1, aux, aux, -1));
            RUN(run_reshape(R_Z_PENCILS_TO_BRICKS, aux, out, aux2));
        } else {
            RUN(run_fft(1, aux, out, -1));
        }
    } else {
        if (is_final_reshape_enabled_) {
            RUN(run_reshape(R_Z_BRICKS_TO_PENCILS, in, aux, aux2));
            RUN(run_fft_transpose(1, aux, aux, +1, T_Y_TO_X, in, aux2));
        } else {
            RUN(run_fft_transpose(1, in, aux, +1, T_Y_TO_X, in, aux2));
        }
        RUN(run_fft(0, in, aux, +1));
        RUN(run_reshape(R_X_PENCILS_TO_BRICKS, aux, out, aux2));
    }
    return DTFFT_SUCCESS;
}

int Plan::execute_z_slab(void* in, void* out, bool fwd, void* aux, bool inplace, void* aux2) {  // :974-1016
    if (is_transpose_plan_) {
        if (inplace) return execute_generic(in, out, fwd, aux, aux2);
        return run_transpose(fwd ? T_X_TO_Z : T_Z_TO_X, in, out, aux);
    }
    if (fwd) {
        RUN(run_fft_transpose(0, in, aux, -1, T_X_TO_Z, out, aux2));
        RUN(run_fft(2, out, out, -1));
    } else {
        RUN(run_fft_transpose(2, in, in, +1, T_Z_TO_X, aux, aux2));
        RUN(run_fft(0, aux, out, +1));
    }
    return DTFFT_SUCCESS;
}

int Plan::execute_z_slab_reshape(void* in, void* out, bool fwd, void* aux, void* aux2) {  // :1018-1055
    if (is_transpose_plan_) {
        if (fwd) {
            RUN(run_reshape(R_X_BRICKS_TO_PENCILS, in, aux, aux2));
            RUN(run_transpose(T_X_TO_Z, aux, out, aux2));
        } else {
            RUN(run_transpose(T_Z_TO_X, in, aux, aux2));
            RUN(run_reshape(R_X_PENCILS_TO_BRICKS, aux, out, aux2));
        }
        return DTFFT_SUCCESS;
    }
    if (fwd) {
        RUN(run_reshape(R_X_BRICKS_TO_PENCILS, in, aux, aux2));
        RUN(run_fft_transpose(0, aux, in, -1, T_X_TO_Z, aux, aux2));
        RUN(run_fft(2, aux, out, -1));
    } else {
        RUN(run_fft_transpose(2, in, aux, +1, T_Z_TO_X, in, aux2));
        RUN(run_fft(0, in, aux, +1));
        RUN(run_reshape(R_X_PENCILS_TO_BRICKS, aux, out, aux2));
    }
    return DTFFT_SUCCESS;
}

int Plan::execute_generic(void* in, void* out, bool fwd, void* aux, void* aux2) {  // :1057-1101
    if (is_transpose_plan_) {
        if (fwd)
            RUN(run_transpose_pair(T_X_TO_Y, in, aux, T_Y_TO_Z, out, aux2));
        else
            RUN(run_transpose_pair(T_Z_TO_Y, in, aux, T_Y_TO_X, out, aux2));
        return DTFFT_SUCCESS;
    }
    if (fwd) {
        RUN(run_fft_transpose(0, in, aux, -1, T_X_TO_Y, out, aux2));
        RUN(run_fft_transpose(1, out, out, -1, T_Y_TO_Z, aux, aux2));
        RUN(run_fft(2, aux, out, -1));
    } else {
        auto zy = handles_.find(T_Z_TO_Y);
        // (not for in-place calls: the last transform would become `in -> in`, which an out-of-place
        // c2r plan cannot do; every rank calls alike, so the decision stays collective)
        if (overlap_chunks_ > 1 && in != out && zy != handles_.end() && zy->second->can_chunk()) {
            // Stage overlap: the reference's choreography (below) transposes back INTO the buffer the
            // Z transform reads, which forbids storing chunk k while chunk k+1 is still being
            // transformed.  `in` is scratch on the backward pass anyway (the reference overwrites it
            // too), so transform it in place and ping-pong in -> aux -> in instead; results are identical.
            RUN(run_fft_transpose(2, in, in, +1, T_Z_TO_Y, aux, aux2));
            RUN(run_fft_transpose(1, aux, aux, +1, T_Y_TO_X, in, aux2));
            RUN(run_fft(0, in, out, +1));
            return DTFFT_SUCCESS;
        }
        RUN(run_fft_transpose(2, in, aux, +1, T_Z_TO_Y, in, aux2));
        RUN(run_fft_transpose(1, in, in, +1, T_Y_TO_X, aux, aux2));
        RUN(run_fft(0, aux, out, +1));
    }
    return DTFFT_SUCCESS;
}

int Plan::execute_generic_reshape(void* in, void* out, bool fwd, void* aux, void* aux2) {  // :1103-1167
    if (is_transpose_plan_) {
        if (fwd) {
            RUN(run_reshape(R_X_BRICKS_TO_PENCILS, in, aux, aux2));
            RUN(run_transpose(T_X_TO_Y, aux, in, aux2));
            if (is_final_reshape_enabled_) {
                RUN(run_transpose(T_Y_TO_Z, in, aux, aux2));
                RUN(run_reshape(R_Z_PENCILS_TO_BRICKS, aux, out, aux2));
            } else {
                RUN(run_transpose(T_Y_TO_Z, in, out, aux2));
            }
        } else {
            if (is_final_reshape_enabled_) {
                RUN(run_reshape(R_Z_BRICKS_TO_PENCILS, in, aux, aux2));
                RUN(run_transpose(T_Z_TO_Y, aux, out, aux2));
            } else {
                RUN(run_transpose(T_Z_TO_Y, in, out, aux2));
            }
            RUN(run_transpose(T_Y_TO_X, out, aux, aux2));
            RUN(run_reshape(R_X_PENCILS_TO_BRICKS, aux, out, aux2));
        }
        return DTFFT_SUCCESS;
    }
    if (fwd) {
        RUN(run_reshape(R_X_BRICKS_TO_PENCILS, in, aux, aux2));
        RUN(run_fft_transpose(0, aux, out, -1, T_X_TO_Y, aux, aux2));
        RUN(run_fft_transpose(1, aux, aux, -1, T_Y_TO_Z, out, aux2));
        if (is_final_reshape_enabled_) {
            RUN(run_fft(2, out, aux, -1));
            RUN(run_reshape(R_Z_PENCILS_TO_BRICKS, aux, out, aux2));
        } else {
            RUN(run_fft(2, out, out, -1));
        }
    } else {
        if (is_final_reshape_enabled_) {
            RUN(run_reshape(R_Z_BRICKS_TO_PENCILS, in, aux, aux2));
            RUN(run_fft_transpose(2, aux, in, +1, T_Z_TO_Y, aux, aux2));
        } else {
            RUN(run_fft_transpose(2, in, in, +1, T_Z_TO_Y, aux, aux2));
        }
        RUN(run_fft_transpose(1, aux, aux, +1, T_Y_TO_X, in, aux2));
        RUN(run_fft(0, in, aux, +1));
        RUN(run_reshape(R_X_PENCILS_TO_BRICKS, aux, out, aux2));
    }
    return DTFFT_SUCCESS;
}
#undef RUN

int Plan::report() const {
    // dtfft_plan.F90:1556-1631, line for line (WRITE_REPORT: rank 0, prefix "dtFFT: "); the lines after
    // "Reshape Backend" are additions of this implementation.
    if (!created_) return DTFFT_ERROR_PLAN_NOT_CREATED;
    if (comm_.rank() != 0) return DTFFT_SUCCESS;
    auto grid_str = [&](const int32_t* d) {
        char buf[64];
        if (ndims_ == 2)
            snprintf(buf, sizeof buf, "%dx%d", d[0], d[1]);
        else
            snprintf(buf, sizeof buf, "%dx%dx%d", d[0], d[1], d[2]);
        return std::string(buf);
    };
    auto line = [](const char* fmt, ...) {
        va_list ap;
        va_start(ap, fmt);
        fputs("dtFFT: ", stdout);
        vfprintf(stdout, fmt, ap);
        fputc('\n', stdout);
        va_end(ap);
    };
    line("**Plan report**");
    line("  dtFFT Version        :  %d.%d.%d", DTFFT_VERSION_MAJOR, DTFFT_VERSION_MINOR, DTFFT_VERSION_PATCH);
    line("  Number of dimensions :  %d", ndims_);
    line("  Global dimensions    :  %s", grid_str(user_dims_).c_str());
    line("  Grid decomposition   :  %s", grid_str(comm_dims_).c_str());
    if (is_reshape_enabled_) {
        // init_grid = brick grid; final_grid = (P1, P2 / c, c) resp. (P1 / c, c), c = bricks pooled along the
        // last axis (src/dtfft_reshape_plan.F90:155-158, 199-204)
        const int last = ndims_ - 1, c = std::max(1, brick_grid_[last]);
        int32_t fin[3] = {1, 1, 1};
        if (ndims_ == 3)
            fin[0] = comm_dims_[1], fin[1] = comm_dims_[2] / c, fin[2] = c;
        else
            fin[0] = comm_dims_[1] / c, fin[1] = c;
        line("  Initial grid         :  %s", grid_str(brick_grid_).c_str());
        line("  Final grid           :  %s", grid_str(fin).c_str());
        line("  Final reshape enabled:  %s", is_final_reshape_enabled_ ? "True" : "False");
    }
    line("  Execution platform   :  CUDA");
    line("  Plan type            :  %s", kind_ == PLAN_C2C ? "Complex-to-Complex" : kind_ == PLAN_R2C ? "Real-to-Complex" : "Real-to-Real");
    line("  Plan precision       :  %s", dtfft_get_precision_string((dtfft_precision_t)precision_));
    line("  FFT Executor type    :  %s", dtfft_get_executor_string((dtfft_executor_t)executor_));
    if (ndims_ == 3) {
        line("  Z-slab enabled       :  %s", is_z_slab_ ? "True" : "False");
        line("  Y-slab enabled       :  %s", is_y_slab_ ? "True" : "False");
    }
    line("  Backend              :  %s", dtfft_get_backend_string((dtfft_backend_t)backend_));
    if (is_reshape_enabled_) line("  Reshape Backend      :  %s", dtfft_get_backend_string((dtfft_backend_t)reshape_backend_));
    line("  Alloc / aux bytes    :  %zu / %zu", alloc_bytes(), aux_bytes());
    line("  Stage overlap chunks :  %d", overlap_chunks_);
    line("  CUDA graph replay    :  %s", graphs_usable() ? "True" : "False");
    {  // how every transposition moves its data on this rank (Plan::exchange_form)
        static const char* const tnames[7] = {"Z_TO_X", "Z_TO_Y", "Y_TO_X", "", "X_TO_Y", "Y_TO_Z", "X_TO_Z"};
        static const char* const fnames[5] = {"local kernel", "NCCL pack / send-recv / unpack", "direct-store kernel",
                                              "copy engines", "direct-store kernel alone, copy engines in pair pipelines"};
        for (const auto& kv : handles_) {
            int f = 0, n = 0;
            if (exchange_form(kv.first, &f, &n) == DTFFT_SUCCESS && kv.first >= -3 && kv.first <= 3 && kv.first != 0)
                line("  Transpose %-6s     :  %s", tnames[kv.first + 3], fnames[f]);
        }
        if (stat_fallbacks_ > 0) line("  NCCL stand-in calls  :  %lld", (long long)stat_fallbacks_);
    }
    line("**End of report**");
    fflush(stdout);
    return DTFFT_SUCCESS;
}

int Plan::destroy() {
    // dtfft_plan.F90:1169-1250
    if (stream_) cudaStreamSynchronize(stream_);
    drop_graphs();
    if (xfer_stream_) {
        cudaStreamSynchronize(xfer_stream_);
        cudaStreamDestroy(xfer_stream_);
        xfer_stream_ = nullptr;
    }
    for (cudaEvent_t e : chunk_events_) cudaEventDestroy(e);
    chunk_events_.clear();
    if (xfer_done_) cudaEventDestroy(xfer_done_);
    xfer_done_ = nullptr;
    handles_.clear();
    rhandles_.clear();
    for (auto& f : fft_) f.reset();
    if (is_aux_alloc_ && aux_ptr_) mem_free(aux_ptr_);
    aux_ptr_ = nullptr;
    is_aux_alloc_ = false;
    while (!allocs_.empty()) mem_free(allocs_.back().ptr);
    peers_.destroy();
    if (nccl_) ncclCommDestroy(nccl_);
    nccl_ = nullptr;
    if (own_stream_ && stream_) cudaStreamDestroy(stream_);
    stream_ = nullptr;
    own_stream_ = false;
    created_ = false;
    cudaGetLastError();
    return DTFFT_SUCCESS;
}

}  // namespace dtfftb
