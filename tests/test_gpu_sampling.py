"""GPU parity: stratified / hierarchical sampling kernels vs the oracle and the reference goldens."""
import pytest
import torch

import fixtures as fx
import ibl_nerf_b200 as ib
from oracle import iblnerf_oracle as orc
from util import G, close, close_mostly

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("s,lindisp,jitter", [(64, False, True), (64, False, False), (192, True, True), (7, False, True)])
def test_stratified_z_bit_exact(s, lindisp, jitter):
    n = 333
    g = torch.Generator().manual_seed(2)
    near = 0.3 + torch.rand(n, 1, generator=g)
    far = 4.0 + 4 * torch.rand(n, 1, generator=g)
    t = torch.rand(n, s, generator=g) if jitter else None
    want = orc.stratified_z(near, far, s, t, lindisp)
    got = ib.ops.stratified_z(near.to(DEV), far.to(DEV), s, None if t is None else t.to(DEV), lindisp)
    assert torch.equal(got.cpu(), want)


def test_inverse_cdf_indices_bit_exact_golden():
    g = G("sample_pdf.npz", DEV)
    inds, samples = ib.ops.inverse_cdf(g["cdf"], g["bins"], g["u"])
    assert torch.equal(inds, g["inds"])                                  # north-star criterion 1
    _, want = orc.inverse_cdf(g["cdf"].cpu(), g["bins"].cpu(), g["u"].cpu())
    assert torch.equal(samples.cpu(), want)                              # same ops, un-contracted -> identical


def test_sample_pdf_golden_and_api():
    g = G("sample_pdf.npz", DEV)
    close_mostly(ib.sample_pdf(g["bins"], g["weights"], 128, det=True), g["s_det"], rtol=1e-5, atol=2e-6, name="det")
    close_mostly(ib.sample_pdf(g["bins"], g["weights"], 128, det=False, pytest=True), g["s_rand_pytest"], rtol=1e-5, atol=2e-6, name="pytest")
    # strided view input (weights[..., 1:-1]) without a copy
    w_full = torch.rand(40, 64, device=DEV)
    z = fx.make_sorted_z(40, 64).to(DEV)
    mids = .5 * (z[:, 1:] + z[:, :-1])
    u = torch.rand(40, 128, device=DEV)
    got = ib.ops.sample_pdf_u(mids, w_full[:, 1:-1], u)
    want = orc.sample_pdf(mids.cpu(), w_full[:, 1:-1].cpu(), u.cpu())
    close_mostly(got, want, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("n,s0,s1", [(1, 64, 128), (257, 64, 128), (33, 192, 384), (5, 8, 3)])
def test_hierarchical_sample_and_merge(n, s0, s1):
    g = torch.Generator().manual_seed(n)
    z = fx.make_sorted_z(n, s0, seed=n)
    w = torch.rand(n, s0, generator=g) ** 3
    u = torch.rand(n, s1, generator=g)
    zs, zm = ib.ops.hierarchical_sample(z.to(DEV), w.to(DEV), u.to(DEV))
    want_s = orc.sample_pdf(.5 * (z[:, 1:] + z[:, :-1]), w[:, 1:-1], u)
    close_mostly(zs, want_s, rtol=1e-5, atol=2e-6, name="z_samples")
    assert torch.equal(zm, torch.sort(torch.cat([z.to(DEV), zs], -1), -1)[0])     # merge is exact given the samples
    assert torch.equal(ib.ops.merge_sort_z(z.to(DEV), zs), zm)


@pytest.mark.parametrize("n,nbins,nsamp", [(301, 191, 384), (77, 511, 1024), (130, 100, 77), (9, 65, 129), (40, 700, 300)])
def test_sample_pdf_longer_rays_vs_oracle(n, nbins, nsamp):
    """The weights -> samples path beyond the shipped 63 / 128 shape (BASELINE config 5 sweep: 191 / 384, 511 / 1024; ragged
    sizes; > 512 bins falls back to the generic kernel) against the oracle's sample_pdf."""
    g = torch.Generator().manual_seed(nbins)
    bins = torch.sort(torch.rand(n, nbins, generator=g) * 7.5 + 0.5, -1)[0]
    w = torch.rand(n, nbins - 1, generator=g) ** 3
    w[0] = 0.0
    u = torch.rand(n, nsamp, generator=g)
    got = ib.ops.sample_pdf_u(bins.to(DEV), w.to(DEV), u.to(DEV))
    close_mostly(got, orc.sample_pdf(bins, w, u), rtol=1e-5, atol=2e-6, name="samples")


def test_sample_pdf_large_properties():
    """BASELINE config 5 scale: monotone in u, inside the bin range, deterministic."""
    n = 1 << 17
    z = torch.sort(torch.rand(n, 64, device=DEV) * 7.5 + 0.5, -1)[0]
    mids = .5 * (z[:, 1:] + z[:, :-1])
    w = torch.rand(n, 62, device=DEV)
    u = torch.sort(torch.rand(n, 128, device=DEV), -1)[0]
    s1 = ib.ops.sample_pdf_u(mids, w, u)
    s2 = ib.ops.sample_pdf_u(mids, w, u)
    assert torch.equal(s1, s2)
    assert (s1[:, 1:] >= s1[:, :-1] - 1e-5).all()
    assert (s1 >= mids[:, :1] - 1e-5).all() and (s1 <= mids[:, -1:] + 1e-5).all()


@pytest.mark.parametrize("nbins,nsamp", [(63, 128), (191, 384), (511, 1024), (2, 5), (64, 33)])
def test_inverse_cdf_bit_exact_vs_searchsorted(nbins, nsamp):
    """Guide-table search == torch.searchsorted(cdf, u, right=True) bit-exact (north-star criterion), for skewed CDFs
    (many entries inside one guide cell), repeated entries, and u hitting entries, 0 and 1 exactly."""
    n = 257
    g = torch.Generator().manual_seed(nbins)
    w = torch.rand(n, nbins - 1, generator=g) ** 8                      # a few dominant bins, the rest ~1e-5 floor
    w[::7] = 0.0                                                         # all-floor rays (uniform CDF)
    pdf = (w + 1e-5) / (w + 1e-5).sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros(n, 1), torch.cumsum(pdf, -1)], -1)
    cdf[1::5, nbins // 2:] = cdf[1::5, nbins // 2:nbins // 2 + 1]        # plateaus: repeated entries
    bins = torch.sort(torch.rand(n, nbins, generator=g) * 7 + 0.5, -1)[0]
    u = torch.rand(n, nsamp, generator=g)
    u[:, 0] = 0.0
    u[:, -1] = 1.0
    k = min(nsamp - 2, nbins)
    u[:, 1:1 + k] = cdf[:, :k]                                           # exactly on entries
    inds, samples = ib.ops.inverse_cdf(cdf.to(DEV), bins.to(DEV), u.to(DEV))
    want = torch.searchsorted(cdf.contiguous(), u.contiguous(), right=True)
    assert torch.equal(inds.cpu(), want)
    _, want_s = orc.inverse_cdf(cdf, bins, u)
    assert torch.equal(samples.cpu(), want_s)
