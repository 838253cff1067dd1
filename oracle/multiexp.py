"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's `multiexp`.

Follows /root/reference/bellman/src/multiexp.rs line by line:
  Source::{next,skip} for (Arc<Vec<G>>, usize) ........ multiexp.rs:53-86
  FullDensity / DensityTracker ........................ multiexp.rs:88-157
  multiexp_inner (one region per window, buckets,
      summation by parts, top-down Horner fold) ....... multiexp.rs:159-250
  multiexp (window rule c, density length assert) ..... multiexp.rs:254-281

Generic over any group object from oracle.curves (G1, G2, Dummy), exactly as the
reference is generic over `PrimeCurve`.
"""
from __future__ import annotations

import math


class SynthesisError(Exception):
    """lib.rs:355-370 -- only the variants this path can produce."""


class UnexpectedIdentity(SynthesisError):
    pass


class UnexpectedEof(SynthesisError):      # SynthesisError::IoError(UnexpectedEof)
    pass


class PolynomialDegreeTooLarge(SynthesisError):
    pass


class Source:
    """multiexp.rs:53-86: cursor over (bases, start index)."""

    def __init__(self, group, bases, idx):
        self.group, self.bases, self.idx = group, bases, idx

    def next(self):
        if len(self.bases) <= self.idx:                       # :55-61
            raise UnexpectedEof("expected more bases from source")
        if self.group.is_identity(self.bases[self.idx]):      # :63-65
            raise UnexpectedIdentity()
        ret = self.bases[self.idx]
        self.idx += 1
        return ret

    def skip(self, amt):
        if len(self.bases) <= self.idx:                       # :74-80
            raise UnexpectedEof("expected more bases from source")
        self.idx += amt


class FullDensity:
    """multiexp.rs:95-114: infinite `true`, no query size."""

    def get_query_size(self):
        return None

    def bits(self, n):
        return [True] * n


class DensityTracker:
    """multiexp.rs:116-157."""

    def __init__(self):
        self.bv = []

    def add_element(self):
        self.bv.append(False)

    def inc(self, idx):
        if not self.bv[idx]:
            self.bv[idx] = True

    def get_total_density(self):
        return sum(self.bv)

    def get_query_size(self):
        return len(self.bv)

    def bits(self, n):
        return self.bv

    def to_words(self):
        """bitvec `BitVec<Lsb0, usize>` raw storage: bit i = bit i%64 of word i//64."""
        words = [0] * ((len(self.bv) + 63) // 64)
        for i, b in enumerate(self.bv):
            if b:
                words[i // 64] |= 1 << (i % 64)
        return words


def window_size(n):
    """multiexp.rs:267-271."""
    if n < 32:
        return 3
    return int(math.ceil(math.log(float(n & 0xFFFFFFFF))))


def _region(group, bases, start, density_bits, exponents, skip, c):
    """The `this` closure, multiexp.rs:173-236 -- one window."""
    acc = group.identity()
    src = Source(group, bases, start)
    buckets = [group.identity()] * ((1 << c) - 1)
    handle_trivial = skip == 0
    for exp, density in zip(exponents, density_bits):          # :191
        if density:
            if exp == 0:                                       # :199-200
                src.skip(1)
            elif exp == 1:                                     # :201-206
                if handle_trivial:
                    acc = group.add(acc, src.next())
                else:
                    src.skip(1)
            else:
                d = (exp >> skip) & ((1 << c) - 1)             # :208-214
                if d != 0:
                    buckets[d - 1] = group.add(buckets[d - 1], src.next())   # :217
                else:
                    src.skip(1)
    running = group.identity()                                 # :229-233
    for b in reversed(buckets):
        running = group.add(running, b)
        acc = group.add(acc, running)
    return acc


def multiexp_inner(group, bases, start, density_bits, exponents, c, num_bits):
    """multiexp.rs:159-250.  Returns the group element or raises the error the
    reference's `try_fold` over `.rev()` would surface (highest failing window,
    first error in scan order inside it)."""
    parts = []
    for skip in range(0, num_bits, c):                         # :238-242
        try:
            parts.append(_region(group, bases, start, density_bits, exponents, skip, c))
        except SynthesisError as e:
            parts.append(e)
    acc = group.identity()
    for part in reversed(parts):                               # :244-249
        if isinstance(part, SynthesisError):
            raise part
        for _ in range(c):
            acc = group.double(acc)
        acc = group.add(acc, part)
    return acc


def multiexp(group, bases, start, density, exponents, num_bits=None):
    """multiexp.rs:254-281.  `bases`/`start` = the (Arc<Vec<G>>, usize) source builder,
    `exponents` = canonical (non-Montgomery) integers."""
    c = window_size(len(exponents))
    qs = density.get_query_size()
    if qs is not None:
        assert qs == len(exponents)                            # :273-278
    if num_bits is None:
        num_bits = group.scalar_field.NUM_BITS
    return multiexp_inner(group, bases, start, density.bits(len(exponents)),
                          exponents, c, num_bits)


def naive(group, bases, exponents):
    """multiexp.rs:299-308 `naive_multiexp` from `test_with_bls12`."""
    acc = group.identity()
    for b, e in zip(bases, exponents):
        acc = group.add(acc, group.mul(b, e))
    return acc


# ---- sharded form: what one shard of a multiexp can know on its own ------------------
FLAG_EOF, FLAG_IDENT_ANY, FLAG_IDENT_TOP = 1, 2, 4      # include/bellman_b200.h: BMPC_MSM_FLAG_*


def shard_flags(group, bases, first_base, density_bits, exponents, n_total, num_bits=None):
    """Error conditions of the exponent slice `exponents` (its first dense position consumes
    bases[first_base]) of a multiexp over n_total exponents, as flag bits.  Derived from
    multiexp.rs:55-65,74-80,191-223: EOF = some dense position finds the cursor at or past the end;
    IDENT_ANY = some window calls next() on an identity base (dense, exponent != 0); IDENT_TOP = the
    reference's HIGHEST window (skip = largest multiple of c below num_bits, c from n_total,
    :238-242,267-271) does.  `flags_status` of the OR over all shards is the status of the whole
    multiexp, because every position below the end of the bases precedes the overrun in scan order."""
    c = window_size(n_total)
    if num_bits is None:
        num_bits = group.scalar_field.NUM_BITS
    top_skip = ((num_bits - 1) // c) * c
    flags, idx = 0, first_base
    for exp, dense in zip(exponents, density_bits):
        if not dense:
            continue
        if idx >= len(bases):
            flags |= FLAG_EOF
        elif exp != 0 and group.is_identity(bases[idx]):
            flags |= FLAG_IDENT_ANY
            if exp == 1:
                top = top_skip == 0                             # next() only in window 0 (:201-206)
            else:
                top = ((exp >> top_skip) & ((1 << c) - 1)) != 0
            if top:
                flags |= FLAG_IDENT_TOP
        idx += 1
    return flags


def flags_status(flags):
    """'eof' | 'identity' | None for the OR of all shards' flags (multiexp.rs:244-249: try_fold over
    the windows from the highest down; EOF fails every window, an identity only those consuming it)."""
    if flags & FLAG_EOF:
        return "identity" if flags & FLAG_IDENT_TOP else "eof"
    return "identity" if flags & FLAG_IDENT_ANY else None
