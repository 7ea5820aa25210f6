// Plan objects and the execute loop of tnc_b200 (C ABI in include/tnc_b200.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>

#include <nvtx3/nvToolsExt.h>     // header-only; ranges cost nothing unless a tool is attached

#include "tnc_internal.h"
#include "tc_gemm.h"

namespace tnc {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return TNC_ERR_CUDA;
}

enum OpKind { OP_LEAVES = 0, OP_EINSUM = 1, OP_PERMUTE = 2, OP_ACCUM = 3 };

struct Op {
    int kind = 0;
    tnc_einsum e{};
    tnc_permute p{};
    tnc_accum a{};
    int leaf_begin = 0, leaf_count = 0;
    int64_t koff_a = -1, koff_b = -1;     // byte offsets into the device blob
    int64_t seg_off = -1;                 // streaming kernel: runs of batches sharing their row of A
    int n_seg = 0;
    int pair_table = -1;                  // OUTER_PAIRS: plan table with the C row block of every (A row, B row) pair
    int chain_len = 0;                    // > 1: head of a run of tiny generic steps executed by one launch; -1: member
    int64_t chain_off = -1;               // the run's ChainStep records in the device blob
    std::shared_ptr<TcGemmOp> tc;         // tensor-core lowering, when algo == TNC_ALGO_TC
    // amax words in the workspace tail (byte offsets, -1: none).  a / b (tensor-core steps): the operand's producer
    // already reduced the operand's largest magnitude there; out (any einsum kernel that can): reduce the largest
    // magnitude of this step's output there, for the tensor-core step that consumes it
    TcAmaxWords amax;
    // slice-id bits (LSB-based) the operation's result depends on; leaf loads and the accumulate: every bit.
    // Only filled in, and only read, with TNC_OPT_SLICE_REUSE
    uint64_t deps = ~0ull;
};

}  // namespace tnc

using namespace tnc;

constexpr int kAmaxWordsPerPhase = 512;      // 2 phases x 512 x 4 bytes = the first 4096 bytes of the tail

struct tnc_plan {
    int dtype = TNC_C64;
    int tc_precision = TNC_TC_3XF16;
    int n_sliced = 0;
    bool finalized = false;
    std::vector<std::vector<int32_t>> tables;
    std::vector<int64_t> table_off;        // byte offsets into the device blob
    std::vector<Op> ops[2];
    std::vector<LeafDev> leaves;
    int64_t leaves_off = 0;
    char* dev_blob = nullptr;
    int64_t workspace_bytes = 0;          // arena + the library's tail (TNC_WORKSPACE_TAIL_BYTES)
    int n_amax_words[2] = {0, 0};         // producer-reduced amax words of each phase (kAmaxWordsPerPhase slots each)
    int64_t amax_words_off(int phase) const { return workspace_bytes - TNC_WORKSPACE_TAIL_BYTES + (int64_t)phase * kAmaxWordsPerPhase * 4; }
    std::atomic<int64_t> last_launches{0};   // written once at the end of an execute / profile call
    // CUDA-graph replay of the slice phase (TNC_OPT_CUDA_GRAPH): one instantiated graph per
    // (workspace, leaf blob, accumulator) the plan has been executed with
    bool use_graph = false;
    bool fuse_amax = true;                // TNC_OPT_FUSE_AMAX
    bool slice_reuse = false;             // TNC_OPT_SLICE_REUSE
    std::atomic<bool> graph_failed{false};
    std::mutex graph_mu;
    struct GraphKey {
        const void *ws, *blob, *out;
        int cls;                          // -1: every operation; t >= 0 (slice reuse): the operations that run when the
                                          // slice-id bits 0 .. t changed
        bool operator<(const GraphKey& o) const {
            return ws != o.ws ? ws < o.ws : blob != o.blob ? blob < o.blob : out != o.out ? out < o.out : cls < o.cls;
        }
    };
    std::map<GraphKey, std::pair<cudaGraphExec_t, int64_t>> graphs;      // exec, launches per replay
    int elem_bytes() const { return 8; }
};

namespace {

bool check_tensor(const tnc_plan* pl, const tnc_tensor& t, const char* what) {
    if (t.rank < 0 || t.rank > TNC_MAX_BITS - 1 || t.rows < 1 || t.offset < 0 || (t.offset & 255)) {
        set_error("%s: bad tensor (offset=%lld rank=%d rows=%d)", what, (long long)t.offset, t.rank, t.rows);
        return false;
    }
    (void)pl;
    return true;
}

bool check_positions(const int8_t* pos, int n, int rank, uint64_t& seen, const char* what) {
    for (int i = 0; i < n; ++i) {
        if (pos[i] < 0 || pos[i] >= rank || ((seen >> pos[i]) & 1ull)) {
            set_error("%s: bit position %d invalid or repeated (rank %d)", what, (int)pos[i], rank);
            return false;
        }
        seen |= 1ull << pos[i];
    }
    return true;
}

int64_t tensor_bytes(const tnc_plan* pl, const tnc_tensor& t) {
    return ((int64_t)t.rows << t.rank) * pl->elem_bytes();
}

int phase_ok(int phase) { return phase == TNC_PHASE_ONCE || phase == TNC_PHASE_SLICE; }

}  // namespace

extern "C" {

int tnc_abi_version(void) { return TNC_ABI_VERSION; }

const char* tnc_last_error(void) { return g_error.c_str(); }

int tnc_plan_create(int32_t dtype, int32_t n_sliced_bonds, tnc_plan** out) {
    if (!out || dtype != TNC_C64 || n_sliced_bonds < 0 || n_sliced_bonds > 63) {
        set_error("plan_create: bad arguments (dtype=%d, sliced bonds=%d)", dtype, n_sliced_bonds);
        return TNC_ERR_INVALID;
    }
    tnc_plan* p = new tnc_plan();
    p->dtype = dtype;
    p->n_sliced = n_sliced_bonds;
    *out = p;
    return TNC_OK;
}

int tnc_plan_set_option(tnc_plan* plan, int32_t option, int64_t value) {
    if (!plan || plan->finalized) {
        set_error("set_option: no plan or plan already finalized");
        return plan ? TNC_ERR_STATE : TNC_ERR_INVALID;
    }
    switch (option) {
        case TNC_OPT_TC_PRECISION:
            if (value != TNC_TC_3XTF32 && value != TNC_TC_3XF16 && value != TNC_TC_F16) {
                set_error("set_option: unknown tensor-core precision %lld", (long long)value);
                return TNC_ERR_INVALID;
            }
            plan->tc_precision = (int)value;
            return TNC_OK;
        case TNC_OPT_CUDA_GRAPH:
            if (value != 0 && value != 1) {
                set_error("set_option: TNC_OPT_CUDA_GRAPH takes 0 or 1, got %lld", (long long)value);
                return TNC_ERR_INVALID;
            }
            plan->use_graph = value != 0;
            return TNC_OK;
        case TNC_OPT_FUSE_AMAX:
            if (value != 0 && value != 1) {
                set_error("set_option: TNC_OPT_FUSE_AMAX takes 0 or 1, got %lld", (long long)value);
                return TNC_ERR_INVALID;
            }
            plan->fuse_amax = value != 0;
            return TNC_OK;
        case TNC_OPT_SLICE_REUSE:
            if (value != 0 && value != 1) {
                set_error("set_option: TNC_OPT_SLICE_REUSE takes 0 or 1, got %lld", (long long)value);
                return TNC_ERR_INVALID;
            }
            plan->slice_reuse = value != 0;
            return TNC_OK;
    }
    set_error("set_option: unknown option %d", option);
    return TNC_ERR_INVALID;
}

void tnc_plan_destroy(tnc_plan* plan) {
    if (!plan) return;
    for (int ph = 0; ph < 2; ++ph)
        for (auto& op : plan->ops[ph]) op.tc.reset();
    for (auto& g : plan->graphs) cudaGraphExecDestroy(g.second.first);
    if (plan->dev_blob) cudaFree(plan->dev_blob);
    delete plan;
}

int tnc_plan_add_table(tnc_plan* plan, const int32_t* data, int64_t n, int32_t* table_id) {
    if (!plan || plan->finalized || !data || n < 0 || !table_id) {
        set_error("add_table: bad arguments or plan already finalized");
        return plan && plan->finalized ? TNC_ERR_STATE : TNC_ERR_INVALID;
    }
    plan->tables.emplace_back(data, data + n);
    *table_id = (int32_t)plan->tables.size() - 1;
    return TNC_OK;
}

int tnc_plan_add_leaves(tnc_plan* plan, int32_t phase, const tnc_leaf* leaves, int32_t n) {
    if (!plan || plan->finalized || !phase_ok(phase) || n < 0 || (n > 0 && !leaves)) {
        set_error("add_leaves: bad arguments");
        return TNC_ERR_INVALID;
    }
    Op op;
    op.kind = OP_LEAVES;
    op.leaf_begin = (int)plan->leaves.size();
    op.leaf_count = n;
    for (int i = 0; i < n; ++i) {
        const tnc_leaf& L = leaves[i];
        if (!check_tensor(plan, L.dst, "leaf")) return TNC_ERR_INVALID;
        if (L.n_sliced < 0 || L.n_sliced > TNC_MAX_SLICED || L.src_rank != L.dst.rank + L.n_sliced ||
            L.src_offset < 0) {
            set_error("leaf %d: inconsistent ranks (src %d, dst %d, sliced %d)", i, L.src_rank, L.dst.rank, L.n_sliced);
            return TNC_ERR_INVALID;
        }
        uint64_t seen = 0;
        if (!check_positions(L.keep_pos, L.dst.rank, L.src_rank, seen, "leaf keep_pos")) return TNC_ERR_INVALID;
        if (!check_positions(L.sliced_pos, L.n_sliced, L.src_rank, seen, "leaf sliced_pos")) return TNC_ERR_INVALID;
        LeafDev d{};
        d.src_offset = L.src_offset;
        d.dst_offset = L.dst.offset;
        d.dst_rank = L.dst.rank;
        d.dst_rows = L.dst.rows;
        d.src_rank = L.src_rank;
        d.n_sliced = L.n_sliced;
        for (int s = 0; s < L.n_sliced; ++s) {
            if (L.sliced_bond[s] < 0 || L.sliced_bond[s] >= plan->n_sliced) {
                set_error("leaf %d: sliced bond index %d out of range", i, (int)L.sliced_bond[s]);
                return TNC_ERR_INVALID;
            }
            if (phase == TNC_PHASE_ONCE) {
                set_error("leaf %d: a sliced leaf cannot be loaded in the ONCE phase", i);
                return TNC_ERR_INVALID;
            }
            d.sliced_pos[s] = L.sliced_pos[s];
            d.sliced_shift[s] = (int8_t)(plan->n_sliced - 1 - L.sliced_bond[s]);
        }
        memcpy(d.keep_pos, L.keep_pos, sizeof(d.keep_pos));
        plan->leaves.push_back(d);
    }
    plan->ops[phase].push_back(op);
    return TNC_OK;
}

int tnc_plan_add_einsum(tnc_plan* plan, int32_t phase, const tnc_einsum* e) {
    if (!plan || plan->finalized || !phase_ok(phase) || !e) {
        set_error("add_einsum: bad arguments");
        return TNC_ERR_INVALID;
    }
    if (!check_tensor(plan, e->a, "einsum A") || !check_tensor(plan, e->b, "einsum B") ||
        !check_tensor(plan, e->c, "einsum C"))
        return TNC_ERR_INVALID;
    if (e->n_m < 0 || e->n_n < 0 || e->n_k < 0 || e->n_h < 0 || e->n_m + e->n_k + e->n_h != e->a.rank ||
        e->n_n + e->n_k + e->n_h != e->b.rank || e->n_m + e->n_n + e->n_h != e->c.rank) {
        set_error("einsum: mode counts (m=%d n=%d k=%d h=%d) do not match ranks (%d, %d, %d)", e->n_m, e->n_n,
                  e->n_k, e->n_h, e->a.rank, e->b.rank, e->c.rank);
        return TNC_ERR_INVALID;
    }
    if (e->n_k > 26) {
        set_error("einsum: %d contracted bits exceed the supported 26", e->n_k);
        return TNC_ERR_UNSUPPORTED;
    }
    uint64_t sa = 0, sb = 0, sc = 0;
    if (!check_positions(e->m_a, e->n_m, e->a.rank, sa, "einsum m_a") ||
        !check_positions(e->k_a, e->n_k, e->a.rank, sa, "einsum k_a") ||
        !check_positions(e->h_a, e->n_h, e->a.rank, sa, "einsum h_a") ||
        !check_positions(e->n_b, e->n_n, e->b.rank, sb, "einsum n_b") ||
        !check_positions(e->k_b, e->n_k, e->b.rank, sb, "einsum k_b") ||
        !check_positions(e->h_b, e->n_h, e->b.rank, sb, "einsum h_b") ||
        !check_positions(e->m_c, e->n_m, e->c.rank, sc, "einsum m_c") ||
        !check_positions(e->n_c, e->n_n, e->c.rank, sc, "einsum n_c") ||
        !check_positions(e->h_c, e->n_h, e->c.rank, sc, "einsum h_c"))
        return TNC_ERR_INVALID;
    if (e->nb != e->c.rows) {
        set_error("einsum: nb (%d) != c.rows (%d)", e->nb, e->c.rows);
        return TNC_ERR_INVALID;
    }
    const int32_t modes[2] = {e->rows_a, e->rows_b};
    const tnc_tensor* ops[2] = {&e->a, &e->b};
    for (int s = 0; s < 2; ++s) {
        const int32_t m = modes[s];
        if (m == TNC_ROWS_NONE) continue;
        if (m == TNC_ROWS_IDENTITY) {
            if (ops[s]->rows < e->nb) {
                set_error("einsum: identity rows but operand has %d rows < nb %d", ops[s]->rows, e->nb);
                return TNC_ERR_INVALID;
            }
            continue;
        }
        if (m < 0 || m >= (int)plan->tables.size() || (int64_t)plan->tables[m].size() < e->nb) {
            set_error("einsum: row table %d missing or shorter than nb=%d", m, e->nb);
            return TNC_ERR_INVALID;
        }
        for (int i = 0; i < e->nb; ++i)
            if (plan->tables[m][i] < 0 || plan->tables[m][i] >= ops[s]->rows) {
                set_error("einsum: row table %d entry %d = %d out of range (rows %d)", m, i, plan->tables[m][i],
                          ops[s]->rows);
                return TNC_ERR_INVALID;
            }
    }
    if (e->flags & TNC_EINSUM_OUTER_ROWS) {
        bool ok = (int64_t)e->a.rows * e->b.rows == e->nb && e->rows_a != TNC_ROWS_NONE && e->rows_b != TNC_ROWS_NONE;
        for (int i = 0; ok && i < e->nb; ++i) {
            const int ra = e->rows_a >= 0 ? plan->tables[e->rows_a][i] : i;
            const int rb = e->rows_b >= 0 ? plan->tables[e->rows_b][i] : i;
            ok = ra == i / e->b.rows && rb == i % e->b.rows;
        }
        if (!ok) {
            set_error("einsum: TNC_EINSUM_OUTER_ROWS set but the rows are not all (A row, B row) pairs, A-major");
            return TNC_ERR_INVALID;
        }
    }
    std::vector<int32_t> pair_rows;
    if (e->flags & TNC_EINSUM_OUTER_PAIRS) {
        bool ok = (int64_t)e->a.rows * e->b.rows == e->nb && e->rows_a != TNC_ROWS_NONE && e->rows_b != TNC_ROWS_NONE &&
                  !(e->flags & TNC_EINSUM_OUTER_ROWS);
        if (ok) pair_rows.assign((size_t)e->nb, -1);
        for (int i = 0; ok && i < e->nb; ++i) {
            const int ra = e->rows_a >= 0 ? plan->tables[e->rows_a][i] : i;
            const int rb = e->rows_b >= 0 ? plan->tables[e->rows_b][i] : i;
            ok = ra < e->a.rows && rb < e->b.rows && pair_rows[(size_t)ra * e->b.rows + rb] < 0;
            if (ok) pair_rows[(size_t)ra * e->b.rows + rb] = i;
        }
        if (!ok) {
            set_error("einsum: TNC_EINSUM_OUTER_PAIRS set but the rows are not every (A row, B row) pair exactly once");
            return TNC_ERR_INVALID;
        }
    }
    if (e->algo != TNC_ALGO_SIMT && e->algo != TNC_ALGO_TC && e->algo != TNC_ALGO_STEM && e->algo != TNC_ALGO_SKINNY) {
        set_error("einsum: unknown algo %d", e->algo);
        return TNC_ERR_INVALID;
    }
    if (e->algo == TNC_ALGO_STEM && !stem_supported(*e, plan->dtype)) {
        set_error("einsum: the streaming kernel does not support this step (k=%d n=%d h=%d, output must be [rows][m][n])",
                  e->n_k, e->n_n, e->n_h);
        return TNC_ERR_UNSUPPORTED;
    }
    if (e->algo == TNC_ALGO_SKINNY && !skinny_supported(*e, plan->dtype, plan->tc_precision)) {
        set_error("einsum: the streaming tensor-core kernel does not support this step (m=%d k=%d n=%d h=%d nb=%d, "
                  "precision %d; needs 2 <= k <= 6 (n <= 6 at k = 6), 1 <= n <= 7, m >= 7, one right operand, output [rows][m][n])",
                  e->n_m, e->n_k, e->n_n, e->n_h, e->nb, plan->tc_precision);
        return TNC_ERR_UNSUPPORTED;
    }
    Op op;
    op.kind = OP_EINSUM;
    op.e = *e;
    if (!pair_rows.empty()) {
        plan->tables.push_back(pair_rows);
        op.pair_table = (int)plan->tables.size() - 1;
    }
    plan->ops[phase].push_back(op);
    return TNC_OK;
}

int tnc_plan_add_permute(tnc_plan* plan, int32_t phase, const tnc_permute* p) {
    if (!plan || plan->finalized || !phase_ok(phase) || !p) {
        set_error("add_permute: bad arguments");
        return TNC_ERR_INVALID;
    }
    if (!check_tensor(plan, p->src, "permute src") || !check_tensor(plan, p->dst, "permute dst")) return TNC_ERR_INVALID;
    if (p->src.rank != p->dst.rank || p->src.rows != p->dst.rows) {
        set_error("permute: src and dst shapes differ");
        return TNC_ERR_INVALID;
    }
    uint64_t seen = 0;
    if (!check_positions(p->perm, p->src.rank, p->src.rank, seen, "permute perm")) return TNC_ERR_INVALID;
    Op op;
    op.kind = OP_PERMUTE;
    op.p = *p;
    plan->ops[phase].push_back(op);
    return TNC_OK;
}

int tnc_plan_add_accum(tnc_plan* plan, int32_t phase, const tnc_accum* a) {
    if (!plan || plan->finalized || !phase_ok(phase) || !a) {
        set_error("add_accum: bad arguments");
        return TNC_ERR_INVALID;
    }
    if (!check_tensor(plan, a->src, "accum src")) return TNC_ERR_INVALID;
    uint64_t seen = 0;
    if (!check_positions(a->out_pos, a->src.rank, a->src.rank, seen, "accum out_pos")) return TNC_ERR_INVALID;
    Op op;
    op.kind = OP_ACCUM;
    op.a = *a;
    plan->ops[phase].push_back(op);
    return TNC_OK;
}

// TNC_OPT_SLICE_REUSE: the slice-id bits behind every slice-phase operation, and the check of what the option asks
// of the caller's layout -- a result that is read by an operation depending on MORE bits is read again in later
// slices without being recomputed, so no other operation of the phase may ever write over it.
static inline uint64_t lowest_bit(uint64_t deps) { return deps & (~deps + 1); }

static int plan_slice_deps(tnc_plan* plan) {
    auto& ops = plan->ops[TNC_PHASE_SLICE];
    struct Range {
        int64_t lo, hi;
    };
    auto range_of = [&](const tnc_tensor& t) { return Range{t.offset, t.offset + tensor_bytes(plan, t)}; };
    std::map<int64_t, uint64_t> at;        // arena offset -> bits behind the tensor last written there
    auto bits_at = [&](int64_t off) {      // ONCE-phase tensors are never written in this phase: no bits
        auto it = at.find(off);
        return it == at.end() ? 0ull : it->second;
    };
    const uint64_t every = plan->n_sliced >= 64 ? ~0ull : ((1ull << plan->n_sliced) - 1);
    for (auto& op : ops) {
        op.deps = every;                   // leaf loads and the accumulate run for every slice
        if (op.kind == OP_LEAVES) {
            for (int i = 0; i < op.leaf_count; ++i) {
                const LeafDev& L = plan->leaves[op.leaf_begin + i];
                uint64_t m = 0;
                for (int q = 0; q < L.n_sliced; ++q) m |= 1ull << L.sliced_shift[q];
                at[L.dst_offset] = m;
            }
        } else if (op.kind == OP_EINSUM) {
            op.deps = bits_at(op.e.a.offset) | bits_at(op.e.b.offset);
            at[op.e.c.offset] = op.deps;
        } else if (op.kind == OP_PERMUTE) {
            op.deps = bits_at(op.p.src.offset);
            at[op.p.dst.offset] = op.deps;
        }
    }
    // the reader of an operation's result: the first later operation that takes a tensor at its offset
    auto reader_of = [&](size_t i, const tnc_tensor& out) -> const Op* {
        for (size_t j = i + 1; j < ops.size(); ++j) {
            const Op& y = ops[j];
            if ((y.kind == OP_EINSUM && (y.e.a.offset == out.offset || y.e.b.offset == out.offset)) ||
                (y.kind == OP_PERMUTE && y.p.src.offset == out.offset) || (y.kind == OP_ACCUM && y.a.src.offset == out.offset))
                return &y;
        }
        return nullptr;
    };
    // TNC_EINSUM_RUN_WITH_READER: such an operation runs whenever its reader does (readers come later: backwards)
    for (size_t i = ops.size(); i-- > 0;) {
        Op& x = ops[i];
        if (x.kind != OP_EINSUM || !(x.e.flags & TNC_EINSUM_RUN_WITH_READER)) continue;
        const Op* y = reader_of(i, x.e.c);
        x.deps = y ? y->deps : every;
    }
    // everything an operation writes
    auto writes_of = [&](const Op& op, std::vector<Range>& out) {
        out.clear();
        if (op.kind == OP_LEAVES)
            for (int i = 0; i < op.leaf_count; ++i) {
                const LeafDev& L = plan->leaves[op.leaf_begin + i];
                out.push_back(Range{L.dst_offset, L.dst_offset + (((int64_t)L.dst_rows << L.dst_rank) * plan->elem_bytes())});
            }
        if (op.kind == OP_EINSUM) {
            out.push_back(range_of(op.e.c));
            if (op.e.algo == TNC_ALGO_TC && op.e.scratch_bytes > 0)
                out.push_back(Range{op.e.scratch_offset, op.e.scratch_offset + op.e.scratch_bytes});
        }
        if (op.kind == OP_PERMUTE) out.push_back(range_of(op.p.dst));
    };
    std::vector<Range> w;
    for (size_t i = 0; i < ops.size(); ++i) {
        const Op& x = ops[i];
        if (x.kind != OP_EINSUM && x.kind != OP_PERMUTE) continue;
        const tnc_tensor& out = x.kind == OP_EINSUM ? x.e.c : x.p.dst;
        // Consecutive slice ids flip the bits 0 .. ctz(s): an operation runs exactly when its LOWEST bit is among them,
        // so two operations run on the same slices iff their lowest bits agree.
        const Op* y = reader_of(i, out);
        const uint64_t reader = y ? y->deps : x.deps;
        if (lowest_bit(reader) == lowest_bit(x.deps)) continue;     // recomputed whenever it is read
        const Range r = range_of(out);
        for (size_t j = 0; j < ops.size(); ++j) {
            if (j == i) continue;
            writes_of(ops[j], w);
            for (const Range& o : w)
                if (o.lo < r.hi && r.lo < o.hi) {
                    set_error("finalize: TNC_OPT_SLICE_REUSE needs the result of slice operation %d (bytes [%lld, %lld)) kept "
                              "intact across slices, but operation %d writes [%lld, %lld)", (int)i, (long long)r.lo,
                              (long long)r.hi, (int)j, (long long)o.lo, (long long)o.hi);
                    return TNC_ERR_INVALID;
                }
        }
    }
    return TNC_OK;
}

int tnc_plan_finalize(tnc_plan* plan, int64_t arena_bytes) {
    if (!plan || plan->finalized || arena_bytes < 0) {
        set_error("finalize: bad arguments or already finalized");
        return plan && plan->finalized ? TNC_ERR_STATE : TNC_ERR_INVALID;
    }
    // every tensor must fit the declared arena; the library's own words (producer-reduced amax words, the
    // slice-id word of graph replay) live in a tail behind it
    const int64_t usable = arena_bytes;
    const int64_t workspace_bytes = ((arena_bytes + 255) & ~(int64_t)255) + TNC_WORKSPACE_TAIL_BYTES;
    auto fits = [&](const tnc_tensor& t) { return t.offset + tensor_bytes(plan, t) <= usable; };
    for (int ph = 0; ph < 2; ++ph)
        for (auto& op : plan->ops[ph]) {
            bool ok = true;
            if (op.kind == OP_EINSUM) ok = fits(op.e.a) && fits(op.e.b) && fits(op.e.c);
            if (op.kind == OP_PERMUTE) ok = fits(op.p.src) && fits(op.p.dst);
            if (op.kind == OP_ACCUM) ok = fits(op.a.src);
            if (!ok) {
                set_error("finalize: an operation addresses memory beyond the declared %lld-byte arena",
                          (long long)arena_bytes);
                return TNC_ERR_NOMEM;
            }
        }
    for (auto& L : plan->leaves)
        if (L.dst_offset + (((int64_t)L.dst_rows << L.dst_rank) * plan->elem_bytes()) > usable) {
            set_error("finalize: a leaf does not fit the declared arena");
            return TNC_ERR_NOMEM;
        }
    // host image of the device blob: row tables, k-offset tables, leaf descriptors
    std::vector<char> blob;
    auto append = [&](const void* data, size_t bytes) {
        size_t off = (blob.size() + 255) & ~(size_t)255;
        blob.resize(off + bytes);
        if (bytes) memcpy(blob.data() + off, data, bytes);
        return (int64_t)off;
    };
    plan->table_off.clear();
    for (auto& t : plan->tables) plan->table_off.push_back(append(t.data(), t.size() * sizeof(int32_t)));
    for (int ph = 0; ph < 2; ++ph)
        for (auto& op : plan->ops[ph]) {
            if (op.kind != OP_EINSUM) continue;
            const tnc_einsum& e = op.e;
            const size_t nk = (size_t)1 << e.n_k;
            std::vector<uint32_t> ka(nk), kb(nk);
            for (size_t k = 0; k < nk; ++k) {
                uint32_t oa = 0, ob = 0;
                for (int i = 0; i < e.n_k; ++i) {
                    const uint32_t bit = (uint32_t)(k >> i) & 1u;
                    oa |= bit << e.k_a[i];
                    ob |= bit << e.k_b[i];
                }
                ka[k] = oa;
                kb[k] = ob;
            }
            op.koff_a = append(ka.data(), nk * sizeof(uint32_t));
            op.koff_b = append(kb.data(), nk * sizeof(uint32_t));
            if (e.algo == TNC_ALGO_STEM && e.rows_a >= 0) {
                const std::vector<int32_t>& ra = plan->tables[e.rows_a];
                std::vector<int32_t> seg;
                for (int32_t i = 0; i < e.nb; ++i)
                    if (i == 0 || ra[i] != ra[i - 1]) seg.push_back(i);
                op.n_seg = (int)seg.size();
                seg.push_back(e.nb);
                op.seg_off = append(seg.data(), seg.size() * sizeof(int32_t));
            }
        }
    if (plan->slice_reuse) {
        if (int rc = plan_slice_deps(plan)) return rc;
    }
    // runs of consecutive tiny generic steps -> one chain launch each (with slice reuse: of steps that depend on the
    // same slice-id bits, so that a chain runs exactly when each of its members must)
    static const bool no_chain = knob("TNC_NO_CHAIN") != nullptr;       // measurement aid
    auto chainable = [&](const Op& op) {
        if (no_chain || plan->dtype != TNC_C64 || op.kind != OP_EINSUM || op.e.algo != TNC_ALGO_SIMT) return false;
        const int64_t total = (int64_t)op.e.nb << op.e.c.rank;
        if (simt_uses_rowdot(op.e.c.rank, op.e.n_k, total, op.e.n_m, op.e.n_n, op.e.n_h)) return false;   // own kernel, own summation order
        return total <= kChainMaxOutputs && (total << op.e.n_k) <= kChainMaxMacs;
    };
    for (int ph = 0; ph < 2; ++ph) {
        auto& ops = plan->ops[ph];
        for (size_t i = 0; i < ops.size();) {
            size_t j = i;
            while (j < ops.size() && chainable(ops[j]) &&
                   (!plan->slice_reuse || lowest_bit(ops[j].deps) == lowest_bit(ops[i].deps))) ++j;
            if (j - i >= 2) {
                std::vector<ChainStep> recs;
                for (size_t t = i; t < j; ++t) {
                    const tnc_einsum& e = ops[t].e;
                    ChainStep c{};
                    c.a_off = e.a.offset;
                    c.b_off = e.b.offset;
                    c.c_off = e.c.offset;
                    c.rows_mode_a = e.rows_a;
                    c.rows_mode_b = e.rows_b;
                    c.rows_a_off = e.rows_a >= 0 ? plan->table_off[e.rows_a] : 0;
                    c.rows_b_off = e.rows_b >= 0 ? plan->table_off[e.rows_b] : 0;
                    c.koff_a_off = ops[t].koff_a;
                    c.koff_b_off = ops[t].koff_b;
                    c.total = (int64_t)e.nb << e.c.rank;
                    c.rank_a = e.a.rank;
                    c.rank_b = e.b.rank;
                    c.rank_c = e.c.rank;
                    c.kb = e.n_k;
                    for (int q = 0; q < TNC_MAX_BITS; ++q) c.c2a[q] = c.c2b[q] = -1;
                    for (int q = 0; q < e.n_m; ++q) c.c2a[e.m_c[q]] = e.m_a[q];
                    for (int q = 0; q < e.n_n; ++q) c.c2b[e.n_c[q]] = e.n_b[q];
                    for (int q = 0; q < e.n_h; ++q) {
                        c.c2a[e.h_c[q]] = e.h_a[q];
                        c.c2b[e.h_c[q]] = e.h_b[q];
                    }
                    recs.push_back(c);
                    ops[t].chain_len = -1;
                }
                ops[i].chain_len = (int)(j - i);
                ops[i].chain_off = append(recs.data(), recs.size() * sizeof(ChainStep));
            }
            i = j > i ? j : i + 1;
        }
    }
    plan->leaves_off = append(plan->leaves.data(), plan->leaves.size() * sizeof(LeafDev));
    if (blob.empty()) blob.resize(256);
    TNC_CUDA(cudaMalloc((void**)&plan->dev_blob, blob.size()));
    TNC_CUDA(cudaMemcpy(plan->dev_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    // tensor-core lowering
    for (int ph = 0; ph < 2; ++ph)
        for (auto& op : plan->ops[ph]) {
            if (op.kind != OP_EINSUM || op.e.algo != TNC_ALGO_TC) continue;
            const int32_t* ra = op.e.rows_a >= 0 ? (const int32_t*)(plan->dev_blob + plan->table_off[op.e.rows_a]) : nullptr;
            const int32_t* rb = op.e.rows_b >= 0 ? (const int32_t*)(plan->dev_blob + plan->table_off[op.e.rows_b]) : nullptr;
            if (op.e.scratch_offset < 0 || op.e.scratch_offset + op.e.scratch_bytes > usable) {
                set_error("finalize: a tensor-core scratch region lies outside the declared arena");
                return TNC_ERR_NOMEM;
            }
            TcGemmOp* tc = nullptr;
            const int32_t* pairs = op.pair_table >= 0 ? (const int32_t*)(plan->dev_blob + plan->table_off[op.pair_table]) : nullptr;
            int rc = tc_gemm_create(op.e, plan->dtype, plan->tc_precision, ra, rb, pairs, &tc);
            if (rc != TNC_OK) return rc;
            op.tc.reset(tc, tc_gemm_destroy);
        }
    plan->workspace_bytes = workspace_bytes;
    // Operand scales without a pass over the operand: when the tensor a tensor-core step (fp16 precisions) reads was
    // written by a streaming or GEMM kernel of the same phase, that kernel reduces the tensor's largest magnitude
    // into a word of the workspace tail as it stores it, and the step's amax launch skips the operand.
    if (plan->tc_precision != TNC_TC_3XTF32 && plan->fuse_amax)
        for (int ph = 0; ph < 2; ++ph) {
            auto& ops = plan->ops[ph];
            int n_words = 0;
            for (size_t i = 0; i < ops.size(); ++i) {
                if (!ops[i].tc) continue;
                for (int which = 0; which < 2; ++which) {
                    const tnc_tensor& x = which ? ops[i].e.b : ops[i].e.a;
                    // the latest earlier operation of the phase that wrote at this offset is the producer (the
                    // buffer is live from there to here); leaves and results of the other phase find none
                    for (size_t j = i; j-- > 0;) {
                        Op& p = ops[j];
                        if (p.kind == OP_PERMUTE && p.p.dst.offset == x.offset) break;
                        if (p.kind != OP_EINSUM || p.e.c.offset != x.offset) continue;
                        const bool same = p.e.c.rank == x.rank && p.e.c.rows == x.rows;
                        const bool emits = p.chain_len == 0 && (p.e.algo == TNC_ALGO_STEM || p.e.algo == TNC_ALGO_SKINNY ||
                                                                (p.tc && tc_gemm_emits_amax(p.tc.get())));
                        if (same && emits && p.amax.out < 0 && n_words < kAmaxWordsPerPhase) {
                            const int64_t off = plan->amax_words_off(ph) + 4 * (int64_t)n_words++;
                            p.amax.out = off;
                            (which ? ops[i].amax.b : ops[i].amax.a) = off;
                        }
                        break;
                    }
                }
            }
            plan->n_amax_words[ph] = n_words;
        }
    plan->finalized = true;
    return TNC_OK;
}

// zeroes the producer-reduced amax words of a phase (before the phase's first operation, once per pass)
static int clear_amax_words(const tnc_plan* plan, int phase, char* ws, cudaStream_t st) {
    if (plan->n_amax_words[phase] == 0) return TNC_OK;
    TNC_CUDA(cudaMemsetAsync(ws + plan->amax_words_off(phase), 0, (size_t)plan->n_amax_words[phase] * 4, st));
    return TNC_OK;
}

int64_t tnc_plan_workspace_bytes(const tnc_plan* plan) { return plan ? plan->workspace_bytes : -1; }

int64_t tnc_plan_num_ops(const tnc_plan* plan, int32_t phase) {
    if (!plan || !phase_ok(phase)) return -1;
    return (int64_t)plan->ops[phase].size();
}

int64_t tnc_plan_last_launches(const tnc_plan* plan) { return plan ? plan->last_launches.load() : -1; }

int64_t tnc_plan_num_fused_amax(const tnc_plan* plan, int32_t phase) {
    if (!plan || !plan->finalized || !phase_ok(phase)) return -1;
    return plan->n_amax_words[phase];
}

// NVTX range per operation ("tc m15 n13 k15 rows1", "skinny ...", "chain x13", "leaves", "accum"), only
// with TNC_NVTX=1: lets ncu / nsys filter and group the launches by step class
struct OpRange {
    bool on;
    OpRange(const Op& op) {
        static const bool enabled = getenv("TNC_NVTX") != nullptr;
        on = enabled;
        if (!on) return;
        char name[96];
        if (op.kind == OP_EINSUM && op.chain_len > 1) snprintf(name, sizeof(name), "chain x%d", op.chain_len);
        else if (op.kind == OP_EINSUM) {
            static const char* algo[] = {"simt", "tc", "stem", "skinny"};
            snprintf(name, sizeof(name), "%s m%d n%d k%d rows%d", algo[op.e.algo & 3], op.e.n_m, op.e.n_n, op.e.n_k, op.e.nb);
        } else snprintf(name, sizeof(name), "%s", op.kind == OP_LEAVES ? "leaves" : op.kind == OP_ACCUM ? "accum" : "permute");
        nvtxRangePushA(name);
    }
    ~OpRange() {
        if (on) nvtxRangePop();
    }
};

// Enqueues one operation; the plan is only read (a finalized plan may be executed from several host
// threads at once, each with its own workspace and stream), launches are counted into *n_launches.
static int run_op(const tnc_plan* plan, const Op& op, const void* leaf_blob, uint64_t slice_id, void* accum_out, char* ws,
                  cudaStream_t st, int64_t* n_launches, LaunchHook hook = nullptr, void* hook_ctx = nullptr,
                  const uint64_t* slice_word = nullptr) {
    if (op.kind == OP_EINSUM && op.chain_len < 0) return TNC_OK;       // ran with the head of its chain
    OpRange range(op);
    switch (op.kind) {
        case OP_LEAVES: {
            *n_launches += op.leaf_count > 0;
            const LeafDev* dl = (const LeafDev*)(plan->dev_blob + plan->leaves_off) + op.leaf_begin;
            return launch_leaf_gather(dl, op.leaf_count, 0, leaf_blob, ws, slice_id, slice_word, plan->dtype, st);
        }
        case OP_EINSUM: {
            const tnc_einsum& e = op.e;
            if (op.chain_len < 0) return TNC_OK;                 // ran with the head of its chain
            if (op.chain_len > 1) {
                *n_launches += 1;
                return launch_simt_chain((const ChainStep*)(plan->dev_blob + op.chain_off), op.chain_len, ws, plan->dev_blob, st);
            }
            if (op.tc) {
                int launches = 0;
                int rc = tc_gemm_run(op.tc.get(), ws, st, hook, hook_ctx, &launches, op.amax);
                *n_launches += launches;
                return rc;
            }
            if (e.algo == TNC_ALGO_STEM || e.algo == TNC_ALGO_SKINNY) {
                const int32_t* ra = e.rows_a >= 0 ? (const int32_t*)(plan->dev_blob + plan->table_off[e.rows_a]) : nullptr;
                const int32_t* rb = e.rows_b >= 0 ? (const int32_t*)(plan->dev_blob + plan->table_off[e.rows_b]) : nullptr;
                *n_launches += 1;
                uint32_t* amax_out = op.amax.out >= 0 ? (uint32_t*)(ws + op.amax.out) : nullptr;
                if (e.algo == TNC_ALGO_SKINNY)
                    return launch_skinny(e, plan->tc_precision, ws + e.a.offset, ws + e.b.offset, ws + e.c.offset, ra, rb, st,
                                         amax_out);
                const int32_t* seg = op.seg_off >= 0 ? (const int32_t*)(plan->dev_blob + op.seg_off) : nullptr;
                return launch_stem(e, ws + e.a.offset, ws + e.b.offset, ws + e.c.offset, ra, rb, seg, op.n_seg, st, amax_out);
            }
            SimtEinsumParams p{};
            p.a = ws + e.a.offset;
            p.b = ws + e.b.offset;
            p.c = ws + e.c.offset;
            p.rows_mode_a = e.rows_a;
            p.rows_mode_b = e.rows_b;
            p.rows_a = e.rows_a >= 0 ? (const int32_t*)(plan->dev_blob + plan->table_off[e.rows_a]) : nullptr;
            p.rows_b = e.rows_b >= 0 ? (const int32_t*)(plan->dev_blob + plan->table_off[e.rows_b]) : nullptr;
            p.rank_a = e.a.rank;
            p.rank_b = e.b.rank;
            p.rank_c = e.c.rank;
            p.kb = e.n_k;
            p.total = (int64_t)e.nb << e.c.rank;
            p.koff_a = (const uint32_t*)(plan->dev_blob + op.koff_a);
            p.koff_b = (const uint32_t*)(plan->dev_blob + op.koff_b);
            for (int q = 0; q < TNC_MAX_BITS; ++q) p.c2a[q] = p.c2b[q] = -1;
            for (int i = 0; i < e.n_m; ++i) p.c2a[e.m_c[i]] = e.m_a[i];
            for (int i = 0; i < e.n_n; ++i) p.c2b[e.n_c[i]] = e.n_b[i];
            for (int i = 0; i < e.n_h; ++i) {
                p.c2a[e.h_c[i]] = e.h_a[i];
                p.c2b[e.h_c[i]] = e.h_b[i];
            }
            *n_launches += 1;
            return launch_simt_einsum(p, plan->dtype, st);
        }
        case OP_PERMUTE: {
            *n_launches += 1;
            if (plan->dtype == TNC_C64) {
                PackDesc d{};
                d.rank = op.p.src.rank;
                d.nb = op.p.src.rows;
                d.rows_mode = TNC_ROWS_IDENTITY;
                d.mode = PACK_COPY;
                memcpy(d.src_pos, op.p.perm, sizeof(d.src_pos));
                return launch_pack(d, ws + op.p.src.offset, ws + op.p.dst.offset, nullptr, st);
            }
            PermuteParams p{};
            p.src = ws + op.p.src.offset;
            p.dst = ws + op.p.dst.offset;
            p.rank = op.p.src.rank;
            p.rows = op.p.src.rows;
            memcpy(p.perm, op.p.perm, sizeof(p.perm));
            return launch_permute(p, plan->elem_bytes(), st);
        }
        case OP_ACCUM: {
            *n_launches += 1;
            if (plan->dtype == TNC_C64 && op.a.src.rank >= 8 && op.a.src.rank < 40) {
                // large results: the tiled permutation kernel in read-add-write mode
                PackDesc d{};
                d.rank = op.a.src.rank;
                d.nb = op.a.src.rows;
                d.rows_mode = TNC_ROWS_IDENTITY;
                d.mode = PACK_ACCUM;
                for (int i = 0; i < d.rank; ++i) d.src_pos[op.a.out_pos[i]] = (int8_t)i;
                return launch_pack(d, ws + op.a.src.offset, accum_out, nullptr, st);
            }
            AccumParams p{};
            p.src = ws + op.a.src.offset;
            p.out = accum_out;
            p.rank = op.a.src.rank;
            p.rows = op.a.src.rows;
            memcpy(p.out_pos, op.a.out_pos, sizeof(p.out_pos));
            return launch_accum(p, plan->dtype, st);
        }
    }
    set_error("execute: unknown op kind %d", op.kind);
    return TNC_ERR_INVALID;
}

int tnc_plan_execute(tnc_plan* plan, const void* leaf_blob, uint64_t slice_begin, uint64_t slice_end,
                     void* accum_out, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!plan || !plan->finalized) {
        set_error("execute: plan is not finalized");
        return TNC_ERR_STATE;
    }
    if (!leaf_blob || !accum_out || (!workspace && plan->workspace_bytes > 0)) {
        set_error("execute: null device pointer");
        return TNC_ERR_INVALID;
    }
    if (workspace_bytes < plan->workspace_bytes) {
        set_error("execute: workspace has %lld bytes, plan needs %lld", (long long)workspace_bytes,
                  (long long)plan->workspace_bytes);
        return TNC_ERR_NOMEM;
    }
    const uint64_t n_slices = plan->n_sliced >= 63 ? ~0ull : (1ull << plan->n_sliced);
    if (slice_begin > slice_end || slice_end > n_slices) {
        set_error("execute: slice range [%llu, %llu) outside [0, %llu)", (unsigned long long)slice_begin,
                  (unsigned long long)slice_end, (unsigned long long)n_slices);
        return TNC_ERR_INVALID;
    }
    if ((uintptr_t)workspace & 255) {
        set_error("execute: workspace must be 256-byte aligned");
        return TNC_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    int64_t launches = 0;
    if (slice_begin == slice_end) {
        plan->last_launches = 0;
        return TNC_OK;
    }
    if (int rc = clear_amax_words(plan, TNC_PHASE_ONCE, ws, st)) return rc;
    for (auto& op : plan->ops[TNC_PHASE_ONCE]) {
        int rc = run_op(plan, op, leaf_blob, 0, accum_out, ws, st, &launches);
        if (rc != TNC_OK) return rc;
    }
    uint64_t s = slice_begin;
    // Replay the slice phase as one CUDA graph per slice: the slice id lives in a workspace word that the leaf
    // gather reads and the graph's last node increments.  Captured once per (workspace, leaf blob, accumulator[,
    // class of changed bits]); a capture that fails falls back to plain launches for good.
    uint64_t* const word = (uint64_t*)(ws + plan->workspace_bytes - 256);
    // cls = -1: every operation of the phase; cls = t (slice reuse): what runs when the bits 0 .. t changed
    auto graph_for = [&](int cls, int64_t* per_replay) -> cudaGraphExec_t {
        tnc_plan::GraphKey key{workspace, leaf_blob, accum_out, cls};
        {
            std::lock_guard<std::mutex> lock(plan->graph_mu);
            auto it = plan->graphs.find(key);
            if (it != plan->graphs.end()) {
                *per_replay = it->second.second;
                return it->second.first;
            }
        }
        // captured on a stream of its own (the caller's may be the legacy default stream, which cannot
        // capture); nothing executes during capture, the graph is then launched on the caller's stream
        const uint64_t mask = cls < 0 ? ~0ull : (cls >= 63 ? ~0ull : ((2ull << cls) - 1));
        cudaGraphExec_t exec = nullptr;
        cudaGraph_t graph = nullptr;
        cudaStream_t cap = nullptr;
        int64_t n = 0;
        int rc = TNC_OK;
        if (cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
            if (cls < 0) rc = clear_amax_words(plan, TNC_PHASE_SLICE, ws, cap);
            if (rc == TNC_OK)
                for (auto& op : plan->ops[TNC_PHASE_SLICE]) {
                    if (cls >= 0 && (op.kind == OP_EINSUM || op.kind == OP_PERMUTE) && !(op.deps & mask)) continue;
                    if (cls >= 0 && op.amax.out >= 0 && cudaMemsetAsync(ws + op.amax.out, 0, 4, cap) != cudaSuccess) rc = TNC_ERR_CUDA;
                    if (rc == TNC_OK) rc = run_op(plan, op, leaf_blob, 0, accum_out, ws, cap, &n, nullptr, nullptr, word);
                    if (rc != TNC_OK) break;
                }
            if (rc == TNC_OK) rc = launch_slice_word(word, 1, true, cap);
            ++n;
            const cudaError_t ce = cudaStreamEndCapture(cap, &graph);
            if (rc == TNC_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess)
                exec = nullptr;
            else if (rc != TNC_OK || ce != cudaSuccess)
                exec = nullptr;
            if (graph) cudaGraphDestroy(graph);
        }
        if (cap) cudaStreamDestroy(cap);
        if (!exec) {
            cudaGetLastError();                      // clear the sticky capture error, if any
            plan->graph_failed = true;
            return nullptr;
        }
        std::lock_guard<std::mutex> lock(plan->graph_mu);
        auto ins = plan->graphs.emplace(key, std::make_pair(exec, n));
        if (!ins.second) {                           // another thread was faster
            cudaGraphExecDestroy(exec);
            exec = ins.first->second.first;
            n = ins.first->second.second;
        }
        *per_replay = n;
        return exec;
    };
    if (plan->use_graph && !plan->slice_reuse && !plan->graph_failed && slice_end - slice_begin >= 2) {
        int64_t per_replay = 0;
        if (cudaGraphExec_t exec = graph_for(-1, &per_replay)) {
            int rc = launch_slice_word(word, slice_begin, false, st);
            if (rc != TNC_OK) return rc;
            for (; s < slice_end; ++s) {
                TNC_CUDA(cudaGraphLaunch(exec, st));
                launches += per_replay;
            }
        }
    }
    if (plan->use_graph && plan->slice_reuse && !plan->graph_failed && slice_end - slice_begin >= 3) {
        // slice reuse: the first slice of the call runs every operation (plain launches); slice s > begin flips the
        // bits 0 .. ctz(s) and replays the graph of that class (at most n_sliced of them, captured on first use)
        for (auto& op : plan->ops[TNC_PHASE_SLICE]) {
            if (op.amax.out >= 0) TNC_CUDA(cudaMemsetAsync(ws + op.amax.out, 0, 4, st));
            int rc = run_op(plan, op, leaf_blob, s, accum_out, ws, st, &launches);
            if (rc != TNC_OK) return rc;
        }
        ++s;
        int rc = launch_slice_word(word, s, false, st);
        if (rc != TNC_OK) return rc;
        for (; s < slice_end; ++s) {
            const int cls = __builtin_ctzll(s);
            int64_t per_replay = 0;
            cudaGraphExec_t exec = graph_for(cls, &per_replay);
            if (!exec) break;                        // capture failed: the plain loop below takes over at slice s
            TNC_CUDA(cudaGraphLaunch(exec, st));
            launches += per_replay;
        }
    }
    for (; s < slice_end; ++s) {
        if (!plan->slice_reuse) {
            if (int rc = clear_amax_words(plan, TNC_PHASE_SLICE, ws, st)) return rc;
        }
        // slice reuse: after the first slice of the call, only what depends on a slice-id bit that changed
        const bool all = s == slice_begin || !plan->slice_reuse;      // (s > slice_begin after a failed graph capture: reuse rule)
        const uint64_t changed = s ^ (s - 1);
        for (auto& op : plan->ops[TNC_PHASE_SLICE]) {
            if (!all && (op.kind == OP_EINSUM || op.kind == OP_PERMUTE) && !(op.deps & changed)) continue;
            if (plan->slice_reuse && op.amax.out >= 0) TNC_CUDA(cudaMemsetAsync(ws + op.amax.out, 0, 4, st));
            int rc = run_op(plan, op, leaf_blob, s, accum_out, ws, st, &launches);
            if (rc != TNC_OK) return rc;
        }
    }
    plan->last_launches = launches;
    return TNC_OK;
}

namespace {
struct ProfileCtx {
    cudaStream_t st;
    std::vector<cudaEvent_t> events;
    bool failed = false;
    void mark() {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess || cudaEventRecord(e, st) != cudaSuccess) failed = true;
        else events.push_back(e);
    }
};
void profile_hook(void* ctx) { ((ProfileCtx*)ctx)->mark(); }
}  // namespace

int tnc_plan_profile(tnc_plan* plan, const void* leaf_blob, uint64_t slice_id, void* accum_out, void* workspace,
                     int64_t workspace_bytes, void* stream, float* ms_once, float* ms_slice) {
    if (!plan || !plan->finalized) {
        set_error("profile: plan is not finalized");
        return TNC_ERR_STATE;
    }
    if (!leaf_blob || !accum_out || !workspace || !ms_once || !ms_slice || workspace_bytes < plan->workspace_bytes ||
        ((uintptr_t)workspace & 255)) {
        set_error("profile: bad arguments");
        return TNC_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    int64_t launches = 0;
    float* outs[2] = {ms_once, ms_slice};
    for (int ph = 0; ph < 2; ++ph) {
        const size_t n = plan->ops[ph].size();
        for (size_t i = 0; i < n * TNC_PROFILE_SLOTS; ++i) outs[ph][i] = 0.f;
        ProfileCtx ctx;
        ctx.st = st;
        std::vector<size_t> first(n + 1, 0);      // index of the event that opens operation i
        int rc = clear_amax_words(plan, ph, ws, st);
        ctx.mark();
        for (size_t i = 0; i < n && rc == TNC_OK; ++i) {
            first[i] = ctx.events.size() - 1;
            const size_t before = ctx.events.size();
            rc = run_op(plan, plan->ops[ph][i], leaf_blob, slice_id, accum_out, ws, st, &launches, profile_hook, &ctx);
            if (rc == TNC_OK && ctx.events.size() == before) ctx.mark();   // single-launch operation
            if (ctx.failed && rc == TNC_OK) {
                set_error("profile: could not record an event");
                rc = TNC_ERR_CUDA;
            }
        }
        first[n] = ctx.events.empty() ? 0 : ctx.events.size() - 1;
        cudaError_t se = cudaStreamSynchronize(st);
        if (rc == TNC_OK && se != cudaSuccess) rc = cuda_fail(se, "cudaStreamSynchronize");
        if (rc == TNC_OK)
            for (size_t i = 0; i < n; ++i) {
                float* o = outs[ph] + i * TNC_PROFILE_SLOTS;
                cudaEventElapsedTime(&o[0], ctx.events[first[i]], ctx.events[first[i + 1]]);
                for (size_t k = first[i], slot = 1; k < first[i + 1] && slot < TNC_PROFILE_SLOTS; ++k, ++slot)
                    cudaEventElapsedTime(&o[slot], ctx.events[k], ctx.events[k + 1]);
            }
        for (auto& e : ctx.events) cudaEventDestroy(e);
        if (rc != TNC_OK) return rc;
    }
    plan->last_launches = launches;
    return TNC_OK;
}

int64_t tnc_einsum_tc_scratch_bytes(int32_t dtype, const tnc_einsum* e) {
    if (!e) return 0;
    const int64_t n = tc_gemm_scratch_bytes(*e, dtype);
    return n < 0 ? 0 : n;
}

int tnc_permute_bits(const void* src, void* dst, int32_t rank, int64_t rows, const int8_t* perm,
                     int32_t elem_bytes, void* stream) {
    if (!src || !dst || !perm || rank < 0 || rank >= TNC_MAX_BITS || rows < 1) {
        set_error("permute_bits: bad arguments");
        return TNC_ERR_INVALID;
    }
    uint64_t seen = 0;
    if (!check_positions(perm, rank, rank, seen, "permute_bits perm")) return TNC_ERR_INVALID;
    if (elem_bytes == 8 && rows < ((int64_t)1 << 31)) {
        // complex64: the shared-memory tiled kernel (coalesced on both sides)
        PackDesc d{};
        d.rank = rank;
        d.nb = (int32_t)rows;
        d.rows_mode = TNC_ROWS_IDENTITY;
        d.mode = PACK_COPY;
        memcpy(d.src_pos, perm, rank);
        return launch_pack(d, src, dst, nullptr, (cudaStream_t)stream);
    }
    PermuteParams p{};
    p.src = src;
    p.dst = dst;
    p.rank = rank;
    p.rows = rows;
    memcpy(p.perm, perm, rank);
    return launch_permute(p, elem_bytes, (cudaStream_t)stream);
}

}  // extern "C"
