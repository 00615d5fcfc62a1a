"""Variable-elimination planner: (network, query vars, evidence vars) -> device program.

The reference interleaves planning and arithmetic inside
`BayesNet._variable_elimination` (/root/reference/sorobn/bayes_net.py:739-794):
it works out the relevant and hidden nodes, slices every CPT by the event, then for
each hidden node multiplies the factors that mention it (`pointwise_mul`,
bayes_net.py:253) and sums it out (`CDTAccessor.sum_out`, bayes_net.py:54).

Here the same decisions are taken once per (query vars, evidence vars) pair and
frozen into a flat int32 program; the arithmetic then runs on the device for any
number of evidence rows.  Each program step is one fused
"product of k factors -> sum out one axis" kernel launch:

    out[o, b] = sum_x  prod_i  in_i[ off_i(o) + x * sx_i + evoff_i(b) ]   (, b)

* `o` runs over the output scope (mixed radix, axis 0 fastest), `b` over evidence
  rows.  A factor that depends on the evidence is *batched*: stored `[scope..., B]`
  with the row axis innermost so that a warp reads 32 x 4 consecutive rows of one
  scope entry with 128-bit loads.
* Evidence never materialises sliced tables: a CPT axis that belongs to an
  evidence variable is indexed per row with that row's state code
  (`evoff_i(b) = sum_k ev[col_k, b] * stride_k`).
* Factors that do not depend on evidence stay unbatched and are computed once per
  call by the flat kernel.

Program layout (int32 words) -- parsed by csrc/sbn_api.cu and by
oracle/program_interp.py (the CPU checker used in tests):

    header : MAGIC VERSION mode n_ev n_tables n_slots n_steps Q post_slot post_batched 0 0
    tables : (offset_floats, size) * n_tables         -- into the float table blob
    slots  : (batched, size_per_row) * n_slots        -- scratch buffers
    steps  : kind n_in out_slot n_axes n_elim | cards[n_axes] | ecards[n_elim] |
             per input: is_slot id batched n_ev (col stride card)*n_ev estrides[n_elim] strides[n_axes]

A step sums out `n_elim` variables at once (0 = product only): the reference's
`sum_out(*variables)` (bayes_net.py:54) also takes several.  With `merge_sum_outs=True` the
planner folds a pure sum-out (an elimination whose only factor is the previous product) into
its producer, which saves writing and re-reading the intermediate (-11.7 % HBM bytes on the
benchmark grid).  It is OFF by default: such launches run on the plain kernel today, whose
50 L1 loads per output cost more than the HBM bytes saved (measured 1.59 ms against 0.78 ms
for the two tiled launches replaced); it becomes the default once the tiled kernel takes an
input that spans both tile axes.

`mode` 0 = flat (one evidence row, nothing batched: evidence offsets are uniform),
1 = batched.  The posterior is produced by the last step into `post_slot`
(`[Q]` or `[Q, B]`, unnormalised) and normalised per row by the engine.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

MAGIC = 0x53424E31  # "SBN1"
VERSION = 4
MAX_IN = 8  # factors fused per launch (csrc/sbn_kernels.cuh: SBN_MAX_IN)
MAX_AXES = 20  # output axes per step (SBN_MAX_AXES)
MAX_EV = 8  # evidence axes per input (SBN_MAX_EV)
TILE_EDGE = 5  # largest register-tile edge of sbn_step_tiled
MAX_ELIM = 3  # variables summed out by one launch (SBN_MAX_ELIM)
MAX_Z = 256  # joint states of the variables summed out by one launch
# Largest table (entries) that may keep evidence axes.  Swept on B200: 4096 is best on all three
# benchmark networks (grid 4.21 ms against 4.34 ms without and 4.38 ms at 16384, where a consumer
# ends up gathering from a 62 KB table for every output).
LIFT_MAX = int(os.environ.get("SOROBN_B200_LIFT_MAX", "4096"))
TILED_MAX_IN = 4  # inputs of one launch of the tiled kernel (csrc: kTiledMaxIn)
SLICE_MIN_BYTES = 64 * 1024  # tables of one launch beyond this are laid out for sliced staging (csrc: SBN_SMEM_BUDGET)
PRELOAD_MAX_IN = 3  # the tiled kernel's preload schedule (all operands of a block in registers)
MODE_FLAT, MODE_BATCHED = 0, 1
KIND_FLAT, KIND_BATCHED = 0, 1
HEADER_WORDS = 12


@dataclass
class CompiledNet:
    """Dense, integer-indexed form of a prepared BayesNet (built by
    `BayesNet.prepare()`; the analogue of the pandas housekeeping at
    bayes_net.py:327-371 plus "ship the tables to the device")."""

    names: list  # var id -> node name, topological order (== BayesNet.nodes)
    domains: list  # var id -> sorted list of state values
    parents: list  # var id -> list of parent ids (sorted by name, like the reference)
    cpt: list  # var id -> float64 ndarray, axes [*parents, var]
    card: np.ndarray = None
    index: dict = field(default_factory=dict)  # name -> var id

    def __post_init__(self):
        self.card = np.array([len(d) for d in self.domains], dtype=np.int64)
        self.index = {n: i for i, n in enumerate(self.names)}

    def scope(self, v):
        return (*self.parents[v], v)

    def ancestors(self, v):
        seen = set()
        stack = list(self.parents[v])
        while stack:
            p = stack.pop()
            if p not in seen:
                seen.add(p)
                stack.extend(self.parents[p])
        return seen


@dataclass
class _Factor:
    is_slot: bool
    buf: int  # table index or slot index
    vars: tuple  # free variable ids
    strides: tuple  # element stride per free variable
    ev: tuple  # ((ev_col, stride, card), ...) -- CPTs and tables that kept evidence axes (never batched)
    batched: bool

    @property
    def depends_on_evidence(self):
        return self.batched or bool(self.ev)


@dataclass
class Step:
    kind: int
    inputs: list  # of (factor, strides-per-eliminated-axis, strides-per-out-axis)
    out_id: int  # logical id of the factor produced (slots are assigned afterwards)
    out_vars: tuple  # axis 0 (fastest) first
    cards: tuple
    elims: tuple  # variables summed out by this launch
    ecards: tuple
    out_slot: int = -1

    @property
    def cx(self):
        return int(np.prod(self.ecards, dtype=np.int64)) if self.ecards else 1


@dataclass
class Plan:
    mode: int
    query: tuple  # query var ids, in output (sorted-name) order, slowest axis first
    evidence: tuple  # evidence var ids == evidence columns
    order: list  # elimination order (var ids)
    tables: list  # var ids whose CPTs are shipped, in table order
    slots: list  # (batched, size)
    steps: list
    post_slot: int
    Q: int
    words: np.ndarray = None
    table_blob: np.ndarray = None
    table_blob64: np.ndarray = None
    table_offsets: list = None

    # ---- cost model (DESIGN.md "algorithmic bytes") -------------------------------
    def bytes_per_row(self):
        """Algorithmic HBM bytes per evidence row: every batched step reads each
        batched input once and writes its output once (fp32), plus the evidence
        codes in and the posterior out."""
        total = 0
        for st in self.steps:
            if st.kind != KIND_BATCHED:
                continue
            for f, _, _ in st.inputs:
                if f.batched:
                    total += 4 * int(np.prod([self._card[v] for v in f.vars], dtype=np.int64))
            total += 4 * int(np.prod(st.cards, dtype=np.int64))
        # normalise: read unnormalised posterior, write posterior
        total += 8 * self.Q
        total += len(self.evidence)  # uint8 codes
        return total

    def step_bytes_per_row(self):
        out = []
        for st in self.steps:
            if st.kind != KIND_BATCHED:
                out.append(0)
                continue
            t = 4 * int(np.prod(st.cards, dtype=np.int64))
            for f, _, _ in st.inputs:
                if f.batched:
                    t += 4 * int(np.prod([self._card[v] for v in f.vars], dtype=np.int64))
            out.append(t)
        return out

    def flops_per_row(self):
        total = 0
        for st in self.steps:
            if st.kind != KIND_BATCHED:
                continue
            total += int(np.prod(st.cards, dtype=np.int64)) * st.cx * len(st.inputs)
        return total

    def scratch_floats_per_row(self):
        return sum(size for batched, size in self.slots if batched)

    def max_factor_per_row(self):
        return max([size for batched, size in self.slots if batched] or [0])


def _min_fill_order(scopes, hidden, card):
    """Greedy min-fill (ties: smaller clique, then lower var id).  The reference
    eliminates in set-iteration order (bayes_net.py:779), which is arbitrary and
    does not change the answer; BASELINE.json asks for min-fill."""
    adj = {}
    for sc in scopes:
        for a in sc:
            adj.setdefault(a, set()).update(b for b in sc if b != a)
    for h in hidden:
        adj.setdefault(h, set())
    remaining = sorted(hidden)
    order = []
    while remaining:
        best_key, best_v = None, None
        for v in remaining:
            nb = list(adj[v])
            fill = 0
            for i in range(len(nb)):
                ai = adj[nb[i]]
                for j in range(i + 1, len(nb)):
                    if nb[j] not in ai:
                        fill += 1
            size = 1
            for u in nb:
                size *= int(card[u])
            key = (fill, size, v)
            if best_key is None or key < best_key:
                best_key, best_v = key, v
        v = best_v
        nb = adj.pop(v)
        for a in nb:
            adj[a].discard(v)
            adj[a].update(b for b in nb if b != a)
        remaining.remove(v)
        order.append(v)
    return order


def build_plan(net: CompiledNet, query, evidence, mode=MODE_BATCHED, order=None, max_in=MAX_IN,
               merge_sum_outs=None, lift_evidence=True, allow_empty_query=False, fuse_elims=None) -> Plan:
    """Plan P(query | evidence) for `net`.

    query / evidence are sequences of var ids.  `evidence` fixes the evidence
    *columns*; their values arrive at run time.
    """
    if merge_sum_outs is None:
        merge_sum_outs = os.environ.get("SOROBN_B200_MERGE", "0") == "1"
    if fuse_elims is None:
        fuse_elims = os.environ.get("SOROBN_B200_FUSE", "1") == "1"
    query = tuple(query)
    evidence = tuple(evidence)
    if not query and not allow_empty_query:
        # bayes_net.py:840-841
        raise ValueError("At least one query variable has to be specified")
    if not query and not evidence:
        raise ValueError("nothing to compute: no query variable and no evidence")
    if set(query) & set(evidence):
        # bayes_net.py:843-845
        raise ValueError("A query variable cannot be part of the event")
    if len(set(query)) != len(query) or len(set(evidence)) != len(evidence):
        raise ValueError("duplicate variable in query or event")
    card = net.card
    for v in evidence:
        if card[v] > 255:
            raise ValueError(f"evidence variable {net.names[v]!r} has {card[v]} states; state codes are uint8")

    # bayes_net.py:763-766
    relevant = {*query, *evidence}
    for v in list(relevant):
        relevant |= net.ancestors(v)
    hidden = relevant - set(query) - set(evidence)
    ev_col = {v: i for i, v in enumerate(evidence)}

    # bayes_net.py:768-776 -- one factor per relevant CPT; evidence axes become
    # per-row gathers instead of boolean filters
    tables = sorted(relevant)
    factors = []
    table_arrays = []
    table_axes = []  # variable of every axis of table_arrays[t], outermost first
    for t, v in enumerate(tables):
        scope = net.scope(v)
        # Shipped layout: free axes first (reference order), evidence axes innermost.  Rows of
        # a warp differ only in their evidence codes, so their gathers of one entry then fall
        # into one 32-byte sector / distinct shared-memory banks instead of `stride` apart.
        perm = [i for i, u in enumerate(scope) if u not in ev_col] + [i for i, u in enumerate(scope) if u in ev_col]
        arr = np.ascontiguousarray(np.transpose(net.cpt[v], perm))
        table_arrays.append(arr)
        pscope = [scope[i] for i in perm]
        table_axes.append(list(pscope))
        shape = [int(card[u]) for u in pscope]
        strides = [int(np.prod(shape[i + 1:], dtype=np.int64)) for i in range(len(shape))]
        free = [(u, s) for u, s in zip(pscope, strides) if u not in ev_col]
        ev = tuple((ev_col[u], s, int(card[u])) for u, s in zip(pscope, strides) if u in ev_col)
        if len(ev) > MAX_EV:
            raise ValueError(f"CPT of {net.names[v]!r} has {len(ev)} evidence axes; the kernel supports {MAX_EV}")
        factors.append(_Factor(False, t, tuple(u for u, _ in free), tuple(s for _, s in free), ev, False))

    if order is None:
        order = _min_fill_order([f.vars for f in factors], hidden, card)
    else:
        order = list(order)
        if set(order) != hidden or len(order) != len(hidden):
            raise ValueError("elimination order must be a permutation of the hidden variables")

    steps = []
    next_id = [0]

    def emit(inputs, elim, out_vars, may_lift=True):
        """One fused launch: multiply `inputs`, sum out `elim` (None: product only).
        out_vars is given fastest axis first.  The output gets a logical id; physical
        slots are assigned after the merge pass.

        Deferred evidence instantiation: when every input is a table (CPTs, or tables built
        this way), the product depends on the evidence row only through the few evidence
        columns those tables are indexed by.  It is then computed ONCE, as a table that keeps
        those evidence variables as ordinary (innermost) axes, by an evidence-independent flat
        launch; consumers gather from it like from a CPT.  The reference filters every CPT by
        the event first (bayes_net.py:772-774); filtering after multiplying gives the same
        numbers and turns per-row work into per-call work."""
        elims = () if elim is None else (tuple(elim) if isinstance(elim, (tuple, list)) else (elim,))
        ecards = tuple(int(card[e]) for e in elims)
        dep = any(f.depends_on_evidence for f in inputs)
        batched = dep and mode == MODE_BATCHED
        if len(out_vars) > MAX_AXES:
            raise ValueError(f"a factor over {len(out_vars)} variables exceeds the kernel's {MAX_AXES} axes")
        cards = tuple(int(card[u]) for u in out_vars)
        size = int(np.prod(cards, dtype=np.int64)) if cards else 1
        if size >= 2**31:
            raise ValueError("a factor with >= 2^31 entries per row does not fit the 32-bit scope index")

        lifted_cols = None
        if batched and lift_evidence and may_lift and not any(f.batched for f in inputs):
            cols = []
            for f in inputs:
                for col, _, c in f.ev:
                    if (col, c) not in cols:
                        cols.append((col, c))
            lifted = size * int(np.prod([c for _, c in cols], dtype=np.int64))
            if len(cols) <= MAX_EV and lifted <= LIFT_MAX and len(out_vars) + len(cols) <= MAX_AXES:
                lifted_cols = cols

        ins = []
        out_id = next_id[0]
        next_id[0] += 1
        if lifted_cols is not None:
            ev_vars = [evidence[col] for col, _ in lifted_cols]
            axes = tuple(ev_vars) + tuple(out_vars)  # evidence axes innermost
            axis_cards = tuple(c for _, c in lifted_cols) + cards
            for f in inputs:
                pos = {u: sd for u, sd in zip(f.vars, f.strides)}
                pos.update({evidence[col]: sd for col, sd, _ in f.ev})
                es = tuple(pos.get(e, 0) for e in elims)
                plain = _Factor(f.is_slot, f.buf, f.vars, f.strides, (), False)  # evidence axes are output axes here
                ins.append((plain, es, tuple(pos.get(u, 0) for u in axes)))
            steps.append(Step(KIND_FLAT, ins, out_id, axes, axis_cards, elims, ecards))
            strides, acc = [], 1
            for c in axis_cards:
                strides.append(acc)
                acc *= c
            n_ev = len(lifted_cols)
            ev = tuple((col, strides[k], c) for k, (col, c) in enumerate(lifted_cols))
            return _Factor(True, out_id, tuple(out_vars), tuple(strides[n_ev:]), ev, False)

        for f in inputs:
            pos = {u: s for u, s in zip(f.vars, f.strides)}
            es = tuple(pos.get(e, 0) for e in elims)
            ins.append((f, es, tuple(pos.get(u, 0) for u in out_vars)))
        steps.append(Step(KIND_BATCHED if batched else KIND_FLAT, ins, out_id, tuple(out_vars), cards, elims, ecards))
        out_strides = []
        acc = 1
        for c in cards:
            out_strides.append(acc)
            acc *= c
        return _Factor(True, out_id, tuple(out_vars), tuple(out_strides), (), batched)

    def fsize(f):
        return int(np.prod([card[u] for u in f.vars], dtype=np.int64)) if f.vars else 1

    def axis_order(inputs, out_set):
        """Fastest-first order of the output axes (`_tile_axes` in the kernel docs).

        Axes 0 and 1 span the register tile of csrc/sbn_step_tiled: an input that lacks
        axis 0 is loaded once per tile column, one that lacks both once per tile.  Every
        ordered pair of output axes is scored by the loads per output it implies
        (batched inputs weigh double: they come from L1/L2/HBM, tables from shared
        memory) and the cheapest pair wins; the remaining axes follow the largest
        batched input's own order so that its reads stay sequential."""
        out = sorted(out_set)
        big = max(inputs, key=lambda f: (f.batched, fsize(f)))
        tail = [u for _, u in sorted(zip(big.strides, big.vars)) if u in out_set]
        tail += [u for u in out if u not in big.vars]
        if len(out) < 2:
            return out
        best_key, best = None, None
        for a0 in out:
            t0 = min(int(card[a0]), TILE_EDGE)
            for a1 in out:
                if a1 == a0:
                    continue
                t1 = min(int(card[a1]), TILE_EDGE)
                loads = 0.0
                for f in inputs:
                    n = (t0 if a0 in f.vars else 1) * (t1 if a1 in f.vars else 1)
                    loads += n * (2.0 if f.batched else 1.0)
                key = (loads / (t0 * t1), -(t0 * t1), a0, a1)
                if best_key is None or key < best_key:
                    best_key, best = key, (a0, a1)
        return [best[0], best[1]] + [u for u in tail if u not in best]

    def lifted_size(fs):
        vs = set().union(*[f.vars for f in fs])
        cols = {(col, c) for f in fs for col, _, c in f.ev}
        return int(np.prod([card[u] for u in vs], dtype=np.int64)) * int(np.prod([c for _, c in cols], dtype=np.int64))

    def combine_tables(inputs, limit=TILED_MAX_IN):
        """A launch with more than TILED_MAX_IN factors falls off the tiled kernel.  When the
        surplus is small tables, multiply those together first: a table-only product is an
        evidence-independent flat launch (see `emit`), and the big launch then gathers one
        value where it gathered several."""
        inputs = list(inputs)
        if mode != MODE_BATCHED or not lift_evidence:
            return inputs
        while len(inputs) > limit:
            tabs = [f for f in inputs if not f.batched]
            best = None
            for i in range(len(tabs)):
                for j in range(i + 1, len(tabs)):
                    sz = lifted_size([tabs[i], tabs[j]])
                    n_cols = len({col for f in (tabs[i], tabs[j]) for col, _, _ in f.ev})
                    if sz <= LIFT_MAX and n_cols <= MAX_EV and (best is None or sz < best[0]):
                        best = (sz, tabs[i], tabs[j])
            if best is None:
                break
            _, fa, fb = best
            inputs = [f for f in inputs if f is not fa and f is not fb]
            inputs.append(emit([fa, fb], None, sorted(set(fa.vars) | set(fb.vars))))
        return inputs

    def product_chain(inputs, elim, final_vars=None):
        # a launch that sums out several variables keeps to the preload schedule's 3 inputs
        fused = isinstance(elim, (tuple, list)) and len(elim) > 1
        inputs = combine_tables(inputs, PRELOAD_MAX_IN if fused else TILED_MAX_IN)
        # bayes_net.py:256 reduces pairwise; fuse up to max_in factors per launch and
        # fold the smallest ones first when there are more
        while len(inputs) > max_in:
            inputs.sort(key=fsize)
            head, inputs = inputs[:max_in], inputs[max_in:]
            union = set().union(*[f.vars for f in head])
            inputs.append(emit(head, None, axis_order(head, union)))
        union = set().union(*[f.vars for f in inputs]) if inputs else set()
        if final_vars is not None:
            assert union == set(final_vars), (union, final_vars)
            return emit(inputs, None, list(final_vars), may_lift=False)
        elims = () if elim is None else (tuple(elim) if isinstance(elim, (tuple, list)) else (elim,))
        out_set = union - set(elims)
        return emit(inputs, elims if elims else None, axis_order(inputs, out_set))

    # bayes_net.py:778-786
    # Fused eliminations.  A later variable w of the order joins x's launch when every factor
    # that mentions w is in x's bucket already, or is a table over variables the bucket covers
    # anyway: sum_w sum_x prod(bucket + those tables).  The joint state space is the one the
    # separate launches walk, but the intermediate over w is never written and read back, and
    # the tile axes are chosen for the launch's real output.
    gone = set()
    for k, x in enumerate(order):
        if x in gone:
            continue
        touching = [f for f in factors if x in f.vars]
        factors = [f for f in factors if x not in f.vars]
        elims = [x]
        if fuse_elims and mode == MODE_BATCHED and any(f.batched for f in touching):
            union = set().union(*[f.vars for f in touching])
            z = int(card[x])
            for w in order[k + 1:]:
                if len(elims) >= MAX_ELIM:
                    break
                if w in gone or w not in union or z * int(card[w]) > MAX_Z:
                    continue
                extra = [f for f in factors if w in f.vars]
                if any(f.batched or not set(f.vars) <= union for f in extra):
                    continue
                if extra and len(touching) + len(extra) > TILED_MAX_IN:
                    continue
                touching += extra
                factors = [f for f in factors if w not in f.vars]
                elims.append(w)
                z *= int(card[w])
        gone.update(elims)
        factors.append(product_chain(touching, tuple(elims)))

    # bayes_net.py:788-794: product of what is left; the answer's levels are sorted
    # by name (bayes_net.py:872-873) and rows by state (sort_index, :875)
    q_sorted = tuple(sorted(query, key=lambda v: net.names[v]))
    post = product_chain(factors, None, final_vars=tuple(reversed(q_sorted)))
    Q = fsize(post)

    steps = _merge_sum_outs(steps, merge_sum_outs)
    if mode == MODE_BATCHED:
        if os.environ.get("SOROBN_B200_DFS", "1") == "1":
            steps = _depth_first_order(steps)
        _relayout_big_tables(steps, table_arrays, table_axes, evidence, card)
    slots, post_slot = _assign_slots(steps, post.buf, keep_unbatched=(mode == MODE_BATCHED))

    plan = Plan(mode=mode, query=q_sorted, evidence=evidence, order=list(order), tables=tables,
                slots=slots, steps=steps, post_slot=post_slot, Q=Q)
    plan._card = card
    _serialise(plan, table_arrays)
    return plan


def _relayout_big_tables(steps, table_arrays, table_axes, evidence, card):
    """Lay a big CPT out for its (single) consumer.

    A launch stages its tables in shared memory; a CPT that does not fit (SLICE_MIN_BYTES: the
    engine's 64 KB budget) can still be staged slice by slice when the tiles one CTA walks touch
    a contiguous part of it.  Tiles enumerate the output axes >= 2 (axis 2 fastest), so the table
    is shipped with those axes outermost in the same significance order, then the eliminated
    variables, then the two tile axes, then its evidence axes (innermost, as for every table).
    Every CPT enters exactly one launch, so nobody else sees the new layout."""
    for st in steps:
        if st.kind != KIND_BATCHED:
            continue
        tabs = [k for k, (f, _, _) in enumerate(st.inputs) if not f.is_slot]
        total = sum(table_arrays[st.inputs[k][0].buf].size for k in tabs)
        for f, _, _ in st.inputs:  # tables built by deferred evidence instantiation are staged too
            if f.is_slot and not f.batched:
                total += int(np.prod([card[u] for u in f.vars] + [c for _, _, c in f.ev], dtype=np.int64))
        if total * 4 <= SLICE_MIN_BYTES:
            continue
        rank = {}  # variable -> significance (higher = outer)
        for j, u in enumerate(st.out_vars):
            rank[u] = (0, j) if j < 2 else (2, j)
        for j, u in enumerate(st.elims):
            rank[u] = (1, j)
        for k in tabs:
            f, _, _ = st.inputs[k]
            t = f.buf
            axes = table_axes[t]
            ev_vars = {evidence[col] for col, _, _ in f.ev}
            free = [u for u in axes if u not in ev_vars]
            assert set(free) == set(f.vars) and all(u in rank for u in free)
            new_axes = sorted(free, key=lambda u: rank[u], reverse=True) + [u for u in axes if u in ev_vars]
            if new_axes == axes:
                continue
            table_arrays[t] = np.ascontiguousarray(np.transpose(table_arrays[t], [axes.index(u) for u in new_axes]))
            table_axes[t] = new_axes
            shape = [int(card[u]) for u in new_axes]
            stride = {u: int(np.prod(shape[i + 1:], dtype=np.int64)) for i, u in enumerate(new_axes)}
            col_of = {evidence[col]: col for col, _, _ in f.ev}
            ev = tuple((col_of[u], stride[u], int(card[u])) for u in new_axes if u in ev_vars)
            g = _Factor(False, t, f.vars, tuple(stride[u] for u in f.vars), ev, False)
            st.inputs[k] = (g, tuple(stride.get(e, 0) for e in st.elims), tuple(stride.get(u, 0) for u in st.out_vars))


def _merge_sum_outs(steps, enabled=True):
    """Fold a step whose ONLY input is the output of an earlier step into that step.

    Such a step is a pure sum-out (or, with nothing to eliminate, a re-layout) of a factor
    that was just produced: `sum_j (sum_x prod_i f_i)`.  Summing both variables in the
    producer costs the same multiplies and saves writing the intermediate and reading it
    back (25 KB of the 223 KB per query on the benchmark grid).  Every intermediate is
    consumed exactly once, so the producer's original output is never needed."""
    if not enabled:
        return steps
    merged = []
    producer = {}  # logical id -> index into merged
    for st in steps:
        f0 = st.inputs[0][0]
        if len(st.inputs) == 1 and f0.is_slot and not f0.ev and f0.buf in producer:
            a = merged[producer[f0.buf]]
            z = a.cx * st.cx
            if a.kind == st.kind and len(a.elims) + len(st.elims) <= MAX_ELIM and z <= MAX_Z:
                new_inputs = []
                for f, es, ss in a.inputs:
                    # the producer's own axis -> stride map: its output axes include the evidence
                    # axes a lifted step keeps (those strides come from f.ev, not f.strides)
                    pos = dict(zip(a.out_vars, ss))
                    new_inputs.append((f, es + tuple(pos.get(y, 0) for y in st.elims),
                                       tuple(pos.get(u, 0) for u in st.out_vars)))
                a.inputs = new_inputs
                a.elims = a.elims + st.elims
                a.ecards = a.ecards + st.ecards
                a.out_vars, a.cards = st.out_vars, st.cards
                producer[st.out_id] = producer.pop(f0.buf)
                a.out_id = st.out_id
                continue
        merged.append(st)
        producer[st.out_id] = len(merged) - 1
    return merged


def _depth_first_order(steps):
    """Re-order the launches depth first.

    The steps form a tree (every intermediate has exactly one consumer, the last step produces
    the posterior); the elimination order interleaves its branches.  Executing one branch to the
    end before starting the next keeps the fewest intermediates alive -- the children of a step
    are visited in decreasing (peak - result) order, which is optimal for trees (Sethi-Ullman) --
    so the engine's on-chip segments (csrc/sbn_chain.h) can hold them in shared memory, and a run
    of `frontier <- sum table x frontier` steps becomes contiguous.  Any topological order gives
    the same numbers; slots are assigned afterwards, for THIS order."""
    by_id = {st.out_id: i for i, st in enumerate(steps)}
    children = []
    for st in steps:
        children.append([by_id[f.buf] for f, _, _ in st.inputs if f.is_slot])
    size = [int(np.prod(st.cards, dtype=np.int64)) if st.kind == KIND_BATCHED else 0 for st in steps]
    peak = [0] * len(steps)
    order_of = [None] * len(steps)
    # children always precede their consumer in the incoming list: one forward pass suffices
    for i, st in enumerate(steps):
        kids = sorted(children[i], key=lambda c: (-(peak[c] - size[c]), c))
        # An expanding product (an output several times the size of its siblings' outputs) goes last, so that its
        # consumer follows it immediately: the engine can then run the two as one launch (csrc/sbn_pair.h), and the
        # big intermediate is not held while the other sub-trees are computed.
        if len(kids) > 1 and os.environ.get("SOROBN_B200_BIG_LAST", "1") == "1":
            big = max(kids, key=lambda c: size[c])
            if all(size[big] >= 2 * size[c] for c in kids if c != big):
                kids = [c for c in kids if c != big] + [big]
        held, worst = 0, 0
        for c in kids:
            worst = max(worst, held + peak[c])
            held += size[c]
        peak[i] = max(worst, held + size[i])
        order_of[i] = kids
    out, stack = [], [(len(steps) - 1, 0)]
    seen = set()
    while stack:  # iterative post-order
        node, k = stack.pop()
        if k < len(order_of[node]):
            stack.append((node, k + 1))
            stack.append((order_of[node][k], 0))
        elif node not in seen:
            seen.add(node)
            out.append(steps[node])
    assert len(out) == len(steps), "a step does not feed the posterior"
    return out


def _assign_slots(steps, post_id, keep_unbatched=False):
    """Physical scratch slots by liveness: an output slot is taken before the step's inputs
    are released (a launch never writes a buffer it reads), best fit among the free slots
    of the same kind, and every intermediate dies with its single consumer.

    keep_unbatched: the evidence-independent tables of a batched program are computed ONCE, when
    the program is created (csrc/sbn_api.cu `run_table_steps`), and then read by every run; their
    slots are never recycled (they are a few KB each)."""
    slots = []  # [batched, size, free]
    where = {}  # logical id -> physical slot

    def alloc(batched, size):
        best = None
        for i, (b, sz, free) in enumerate(slots):
            if free and b == batched and sz >= size and (best is None or sz < slots[best][1]):
                best = i
        if best is None:
            slots.append([batched, size, False])
            return len(slots) - 1
        slots[best][2] = False
        return best

    # The batched inputs of a batched step are released one batched step LATE: the engine may run a
    # step and its consumer as ONE launch that reads the first step's operand and writes the second
    # step's output (csrc/sbn_pair.h), so those two must never share a buffer.
    parked = []
    for st in steps:
        size = int(np.prod(st.cards, dtype=np.int64)) if st.cards else 1
        st.out_slot = alloc(st.kind == KIND_BATCHED, size)
        if st.kind == KIND_BATCHED:
            for phys in parked:
                slots[phys][2] = True
            parked = []
        new_inputs = []
        for f, es, ss in st.inputs:
            if f.is_slot:
                phys = where.pop(f.buf)
                if slots[phys][0] and st.kind == KIND_BATCHED:
                    parked.append(phys)
                else:
                    slots[phys][2] = not (keep_unbatched and not slots[phys][0])
                f = _Factor(True, phys, f.vars, f.strides, f.ev, f.batched)
            new_inputs.append((f, es, ss))
        st.inputs = new_inputs
        where[st.out_id] = st.out_slot
    return [(bool(b), int(sz)) for b, sz, _ in slots], where[post_id]


def _serialise(plan: Plan, table_arrays):
    blob = []
    offsets = []
    off = 0
    for arr in table_arrays:
        # plain probabilities: every intermediate entry is then <= 1 and fp32 cannot overflow
        # (DESIGN.md "Precision and fp32 range")
        t = arr.astype(np.float64).reshape(-1)
        pad = (-t.size) % 4  # keep every table 16-byte aligned and sized (bulk-TMA copies)
        offsets.append((off, t.size))
        blob.append(t)
        if pad:
            blob.append(np.zeros(pad, dtype=np.float64))
        off += t.size + pad
    # float64 copy: only the CPU checker (oracle/program_interp.py) reads it, to
    # separate planner errors from fp32 rounding; the device gets the fp32 blob
    plan.table_blob64 = np.concatenate(blob) if blob else np.zeros(0, dtype=np.float64)
    plan.table_blob = plan.table_blob64.astype(np.float32)
    plan.table_offsets = offsets

    w = [MAGIC, VERSION, plan.mode, len(plan.evidence), len(plan.tables), len(plan.slots), len(plan.steps),
         plan.Q, plan.post_slot, int(plan.slots[plan.post_slot][0]), 0, 0]
    assert len(w) == HEADER_WORDS
    for o, s in offsets:
        w += [o, s]
    for b, s in plan.slots:
        w += [int(b), s]
    for st in plan.steps:
        w += [st.kind, len(st.inputs), st.out_slot, len(st.cards), len(st.ecards)]
        w += list(st.cards)
        w += list(st.ecards)
        for f, estrides, strides in st.inputs:
            w += [int(f.is_slot), f.buf, int(f.batched), len(f.ev)]
            for col, s, c in f.ev:
                w += [col, s, c]
            w += list(estrides)
            w += list(strides)
    arr = np.asarray(w, dtype=np.int64)
    if arr.max(initial=0) >= 2**31:
        raise ValueError("program word overflows int32")
    plan.words = arr.astype(np.int32)
