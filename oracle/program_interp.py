"""CPU interpreter of the flat device program (TEST INFRASTRUCTURE, not product).

`sorobn_b200.planner` serialises a variable-elimination plan into int32 words that
`csrc/sbn_api.cu` parses and runs on the GPU.  This module parses the very same
words with numpy and executes them element by element, in float64 (checker) or
float32 (to predict the device's rounding).  Tests use it to check, without a GPU,
that the planner's strides / evidence gathers / slot reuse are right: its output
must equal `oracle.ve_oracle.query` (the restatement of
/root/reference/sorobn/bayes_net.py:739-794) row by row.

Only tests import this; the product never does.
"""
from __future__ import annotations

import numpy as np

MAGIC = 0x53424E31
HEADER_WORDS = 12


def parse(words):
    w = [int(x) for x in np.asarray(words).tolist()]
    assert w[0] == MAGIC, "bad magic"
    hdr = dict(version=w[1], mode=w[2], n_ev=w[3], n_tables=w[4], n_slots=w[5], n_steps=w[6], Q=w[7],
               post_slot=w[8], post_batched=w[9])
    p = HEADER_WORDS
    tables = []
    for _ in range(hdr["n_tables"]):
        tables.append((w[p], w[p + 1]))
        p += 2
    slots = []
    for _ in range(hdr["n_slots"]):
        slots.append((w[p], w[p + 1]))
        p += 2
    steps = []
    for _ in range(hdr["n_steps"]):
        kind, n_in, out_slot, n_axes, n_elim = w[p:p + 5]
        p += 5
        cards = w[p:p + n_axes]
        p += n_axes
        ecards = w[p:p + n_elim]
        p += n_elim
        ins = []
        for _ in range(n_in):
            is_slot, buf, batched, n_ev = w[p:p + 4]
            p += 4
            ev = []
            for _ in range(n_ev):
                ev.append((w[p], w[p + 1], w[p + 2]))
                p += 3
            estrides = w[p:p + n_elim]
            p += n_elim
            strides = w[p:p + n_axes]
            p += n_axes
            ins.append(dict(is_slot=is_slot, buf=buf, batched=batched, estrides=estrides, ev=ev, strides=strides))
        steps.append(dict(kind=kind, out_slot=out_slot, cards=cards, ecards=ecards, inputs=ins))
    assert p == len(w), (p, len(w))
    return hdr, tables, slots, steps


def run(words, table_blob, ev_codes, n_rows=None, dtype=np.float64, return_totals=False):
    """Execute the program.  ev_codes: uint8 array [n_ev, B] (n_rows gives B when
    there are no evidence columns).  Returns the normalised posterior [Q, B]
    (state-major, like the C-ABI's output)."""
    hdr, tables, slots, steps = parse(words)
    ev_codes = np.asarray(ev_codes, dtype=np.uint8)
    if hdr["n_ev"]:
        ev_codes = ev_codes.reshape(hdr["n_ev"], -1)
        B = ev_codes.shape[1]
        assert n_rows is None or n_rows == B
    else:
        B = 1 if n_rows is None else int(n_rows)
    if hdr["mode"] == 0:
        assert B == 1, "flat programs take exactly one evidence row"
    blob = np.asarray(table_blob, dtype=dtype)
    tabs = [blob[o:o + s] for o, s in tables]
    bufs = [None] * len(slots)

    for st in steps:
        cards = st["cards"]
        n_out = int(np.prod(cards, dtype=np.int64)) if cards else 1
        # digits of every output index, axis 0 fastest
        o = np.arange(n_out, dtype=np.int64)
        digits = []
        rem = o.copy()
        for c in cards:
            digits.append(rem % c)
            rem //= c
        batched_out = st["kind"] == 1
        assert all(not (i["is_slot"] and i["buf"] == st["out_slot"]) for i in st["inputs"]), "output aliases an input"
        rows = B if (batched_out or hdr["mode"] == 0) else 1
        acc = np.zeros((n_out, rows), dtype=dtype)
        cx = int(np.prod(st["ecards"], dtype=np.int64)) if st["ecards"] else 1
        for x in range(cx):
            # joint state x of the eliminated variables, first variable fastest
            xd, rem_x = [], x
            for c in st["ecards"]:
                xd.append(rem_x % c)
                rem_x //= c
            prod = np.ones((n_out, rows), dtype=dtype)
            for inp in st["inputs"]:
                off = np.zeros(n_out, dtype=np.int64)
                for d, s in zip(digits, inp["strides"]):
                    off += d * s
                off = off + sum(d * s for d, s in zip(xd, inp["estrides"]))
                evoff = np.zeros(rows, dtype=np.int64)
                for col, s, c in inp["ev"]:
                    evoff = evoff + np.minimum(ev_codes[col, :rows].astype(np.int64), c - 1) * s
                src = bufs[inp["buf"]] if inp["is_slot"] else tabs[inp["buf"]]
                if inp["batched"]:
                    assert inp["is_slot"] and src.ndim == 2 and not inp["ev"]
                    vals = src[off][:, :rows]
                else:
                    flat = src.reshape(-1)
                    vals = flat[off[:, None] + evoff[None, :]]
                prod = (prod * vals).astype(dtype)
            acc = (acc + prod).astype(dtype)
        if batched_out:
            bufs[st["out_slot"]] = acc
        else:
            assert rows == 1
            bufs[st["out_slot"]] = acc.reshape(-1)

    post = bufs[hdr["post_slot"]]
    if post.ndim == 1:
        post = np.repeat(post[:, None], B, axis=1)
    post = post[:hdr["Q"]]
    total = post.sum(axis=0, keepdims=True, dtype=dtype)
    with np.errstate(invalid="ignore", divide="ignore"):
        normalised = (post / total).astype(dtype)
    if return_totals:  # the normaliser is P(event) per row (sbn_program_evidence_host)
        return normalised, total.reshape(-1)
    return normalised
