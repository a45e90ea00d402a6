"""Single-GPU execution of EVERY rank of a sharded denoiser forward, one after the other (test infrastructure).

A first unsharded forward records the projected K|V of all views at every EPPA block (`WarpAttn.kv_tap`). Each rank
(CFG shard bs, view shard vs) then runs the real sharded code path — input slicing, local views only, panorama branch
of its CFG shard, sliced bias / PE tables — with `ReplayParallel` standing in for parallel.ViewParallel: its
`gather_views` checks that the rank's own K|V slice equals what the unsharded run produced for those views and hands
back the recorded K|V of ALL views (what the NCCL all-gather delivers); `gather_outputs` collects the rank's eps
outputs. The driver's 1-GPU box thereby exercises the sharded forward that tests/test_gpu_multi.py needs N GPUs for."""
import torch

from panfusion_b200.eppa import WarpAttn
from panfusion_b200.parallel import ViewParallel


class ReplayParallel(ViewParallel):
    def __init__(self, batch_shards, view_shards, bs, vs, recorded):
        self.batch_shards, self.view_shards = batch_shards, view_shards
        self.world, self.rank = batch_shards * view_shards, bs * view_shards + vs
        self.bs, self.vs = bs, vs
        self.segments = None
        self.device_gather = False
        self.recorded, self.block = recorded, 0
        self.worst_local = 0.0
        self.outputs = None

    def configure(self, b, m):
        pass

    def gather_views(self, x):
        if self.view_shards == 1:
            return x
        full = self.recorded[self.block]            # [b_full, m*P, 2C] of the unsharded run
        self.block += 1
        bl, L, C = x.shape
        mine = full[self.bs * bl:(self.bs + 1) * bl]
        own = mine[:, self.vs * L:(self.vs + 1) * L]
        self.worst_local = max(self.worst_local, (own.float() - x.float()).abs().max().item())
        return mine.contiguous()

    def gather_outputs(self, sample_loc, pano_loc, b, m):
        self.outputs = (sample_loc.clone(), pano_loc.clone())
        bl, ml = b // self.batch_shards, m // self.view_shards
        sample = torch.full((b, m, *sample_loc.shape[2:]), float("nan"), dtype=sample_loc.dtype, device=sample_loc.device)
        pano = torch.full((b, *pano_loc.shape[1:]), float("nan"), dtype=pano_loc.dtype, device=pano_loc.device)
        sample[self.bs * bl:(self.bs + 1) * bl, self.vs * ml:(self.vs + 1) * ml] = sample_loc
        pano[self.bs * bl:(self.bs + 1) * bl] = pano_loc
        return sample, pano


def _drop_identity_caches(model):
    """The text K/V cache is keyed on the identity of the caller's (UNSLICED) prompt tensor — per process that is
    exact, but here every "rank" shares one model object and one prompt tensor while needing a different slice."""
    for br in model._branches[:2]:
        if br is not None:
            br._text_key = None


def run_unsharded_recording(model, inputs):
    rec = []
    WarpAttn.kv_tap = lambda kv: rec.append(kv.clone())
    try:
        model._par = None
        out = model(**inputs)
    finally:
        WarpAttn.kv_tap = None
    return out, rec


def run_all_ranks(model, inputs, batch_shards, view_shards, recorded):
    """-> (sample, pano) assembled from the ranks' own outputs, worst |local K|V - recorded| over ranks."""
    b, m = inputs["latents"].shape[:2]
    sample = pano = None
    worst = 0.0
    try:
        for bs in range(batch_shards):
            for vs in range(view_shards):
                par = ReplayParallel(batch_shards, view_shards, bs, vs, recorded)
                model._par = par
                _drop_identity_caches(model)
                s, p = model(**inputs)
                if sample is None:
                    sample, pano = s.clone(), p.clone()
                else:
                    ok = ~torch.isnan(s)
                    sample[ok] = s[ok]
                    if vs == 0:
                        okp = ~torch.isnan(p)
                        pano[okp] = p[okp]
                    else:  # the panorama branch is replicated over the view shards of a CFG shard: must agree
                        bl = b // batch_shards
                        assert torch.equal(p[bs * bl:(bs + 1) * bl], pano[bs * bl:(bs + 1) * bl])
                worst = max(worst, par.worst_local)
    finally:
        model._par = None
        _drop_identity_caches(model)
    assert not torch.isnan(sample).any() and not torch.isnan(pano).any()
    return sample, pano, worst
