"""Host model of the tiled kernel's box-issuing protocol (csrc/mcs_stitch_tiled.cu: ``issuer_step`` and the frame
loops that call it), run over random chunk sequences.

In the kernel one thread - lane 0 of warp 0 - issues every TMA box of its CTA, one call per unit its own warp
consumes, and the same warp waits for those boxes: a unit whose box is never issued hangs the CTA for good.  The
rules that keep that from happening are few but easy to break (BAND units stage several boxes per frame and only
the BAND frame loop's issuer calls know it; ZERO chunks stage none; the issuer may run a bounded number of chunks
ahead).  This model restates them and checks, for thousands of random sweeps, that every box a consumer waits
for has been issued, in the order the consumers take them, and that the ring is never overrun.  It is a model
of the protocol, not of the arithmetic; the GPU tests hold the kernel itself."""
import random

import pytest

ZERO, COPY, WARP, BAND = 0, 1, 2, 4
LEAD = 10          # TILED_SCHED_LEAD
SLACK = 2          # TILED_LOOKAHEAD_SLACK


class Deadlock(Exception):
    pass


class Model(object):
    def __init__(self, chunks, stages, bands):
        self.chunks = chunks          # dicts: cls, frames, n_ov (BAND), zero_base (BAND)
        self.stages = stages
        self.bands = bands            # the BAND-aware instantiation
        self.k = -1                   # chunk the issuer stands in
        self.f = self.f1 = 0
        self.zero = False
        self.issued = []              # boxes in issue order: (chunk, frame, box index)
        self.nl, self.li, self.li0 = 1, 0, 0
        self.consumed = 0

    # -- issuer_step<BANDS, INBAND> ---------------------------------------------------------------
    def issuer_step(self, inband, k_cons, consumed):
        if len(self.issued) - consumed >= self.stages - 1:
            return
        if self.f == self.f1:
            kn = self.k + 1
            if kn > k_cons + (1 if self.zero else LEAD):
                return
            if kn >= len(self.chunks):
                return
            ch = self.chunks[kn]
            if self.bands and ch["cls"] == BAND and not (inband and kn <= k_cons + 1):
                return
            self.k = kn
            self.zero = ch["cls"] == ZERO
            if self.zero or ch["frames"] == 0:
                return
            self.f, self.f1 = 0, ch["frames"]
            if self.bands and inband:
                self.nl = 1 + ch.get("n_ov", 0)
                self.li0 = 1 if ch.get("zero_base") else 0
                self.li = self.li0
        li = 0
        if self.bands and inband:
            li = self.li
        assert self.chunks[self.k]["cls"] != BAND or (self.bands and inband), \
            "an ordinary issuer call is driving a BAND chunk: its overlay boxes would never be issued"
        self.issued.append((self.k, self.f, li))
        if self.bands and inband:
            li += 1
            if li == self.nl:
                li = self.li0
                self.f += 1
            self.li = li
        else:
            self.f += 1

    # -- the consumers (warp 0: the one that also issues) -----------------------------------------------
    def wait_box(self, expect):
        if self.consumed >= len(self.issued):
            raise Deadlock("waiting for box %r that was never issued (issuer stands in chunk %d, frame %d of %d)"
                           % (expect, self.k, self.f, self.f1))
        assert self.issued[self.consumed] == expect, (self.issued[self.consumed], expect)
        assert len(self.issued) - self.consumed <= self.stages, "ring overrun"
        self.consumed += 1

    def run(self):
        for _ in range(self.stages - SLACK):                     # prologue (BAND-aware: may not run ahead into a BAND chunk 1)
            self.issuer_step(self.bands, -1 if self.bands else 0, -1)
        for k, ch in enumerate(self.chunks):
            cls, n_fr = ch["cls"], ch["frames"]
            if cls == BAND:
                assert self.bands
                for _ in range(self.stages - SLACK):             # refill at BAND chunk entry
                    self.issuer_step(True, k, self.consumed)
                for f in range(n_fr):
                    self.issuer_step(True, k, self.consumed)
                    if not ch.get("zero_base"):
                        self.wait_box((k, f, 0))
                    for o in range(ch["n_ov"]):
                        self.issuer_step(True, k, self.consumed)
                        self.wait_box((k, f, 1 + o))
            else:
                for f in range(n_fr):
                    self.issuer_step(False, k, self.consumed)
                    if cls != ZERO:
                        self.wait_box((k, f, 0))
        assert self.consumed == len(self.issued), "boxes issued that nobody consumes"


def random_sweep(rng, with_bands, spread):
    """Chunks as the launcher makes them: every chunk has at least one frame (mcs_launch_tiled drops the split of
    the last tiles when the last frame block is shorter than the split factor)."""
    n = rng.randint(1, 60)
    chunks = []
    for i in range(n):
        frames = rng.choice([1, 1, 2, 3, 5, 16])
        r = rng.random()
        if with_bands and ((spread and r < 0.25) or (not spread and i < n // 4)):
            n_ov = rng.choice([1, 1, 2])
            chunks.append(dict(cls=BAND, frames=frames, n_ov=n_ov, zero_base=rng.random() < 0.3))
        elif r < 0.5:
            chunks.append(dict(cls=WARP, frames=frames))
        elif r < 0.75:
            chunks.append(dict(cls=COPY, frames=frames))
        else:
            chunks.append(dict(cls=ZERO, frames=frames))
    # several frame blocks: the sweep repeats
    return chunks * rng.choice([1, 1, 2, 3])


@pytest.mark.parametrize("with_bands,spread", [(False, False), (True, False), (True, True)])
def test_every_box_is_issued_before_it_is_waited_for(with_bands, spread):
    rng = random.Random(1234 + 2 * with_bands + spread)
    for _ in range(1500):
        chunks = random_sweep(rng, with_bands, spread)
        stages = rng.choice([3, 4, 6, 7, 8])
        Model(chunks, stages, with_bands).run()


def test_the_model_catches_an_issuer_that_runs_ahead_into_a_later_band_chunk():
    """The hole the entry rule closes: without it (a BAND chunk entered from a BAND frame loop however far ahead it
    is), ordinary frame loops end up driving a BAND chunk."""
    class Loose(Model):
        def issuer_step(self, inband, k_cons, consumed):
            if self.f == self.f1 and self.k + 1 < len(self.chunks) and inband:
                ch = self.chunks[self.k + 1]
                if ch["cls"] == BAND and self.k + 1 > k_cons + 1:
                    # what the rule forbids: pretend the consumers were there
                    return Model.issuer_step(self, inband, self.k + 1, consumed)
            return Model.issuer_step(self, inband, k_cons, consumed)
    chunks = [dict(cls=BAND, frames=4, n_ov=1), dict(cls=WARP, frames=1), dict(cls=WARP, frames=1),
              dict(cls=BAND, frames=4, n_ov=1), dict(cls=WARP, frames=2)]
    Model(chunks, 8, True).run()
    with pytest.raises((AssertionError, Deadlock)):
        Loose(chunks, 8, True).run()


def test_chunks_without_frames_are_why_the_launcher_never_makes_them():
    """A resampled chunk with no frames costs the issuer a call and gives it none back: with a short ring the
    consumers reach the next chunk before its first box was issued, and their single call there only steps over
    the empty chunk."""
    chunks = [dict(cls=WARP, frames=2), dict(cls=WARP, frames=0), dict(cls=WARP, frames=0), dict(cls=WARP, frames=0),
              dict(cls=WARP, frames=1)]
    with pytest.raises(Deadlock):
        Model(chunks, 3, False).run()
