"""TEST INFRASTRUCTURE -- oracle-backed stand-in for ``CudaSeqBackend`` (same interface, torch-CPU math from
``oracle/restate.py``), so that the host logic of ``SequenceParallelSeparator`` (chunk plan, halo exchange, summary
exchange / fold order, decoder seam, assembly) runs on CPU under ``gloo`` at world size 2."""
import torch
import torch.nn.functional as F

from oracle import restate


class OracleSeqBackend:
    def __init__(self, hp, sds):
        self.hp, self.sds = hp, sds
        self.m = sds["masknet"]
        self.n_layers = hp.n_mamba
        self.di = hp.d_inner

    def begin(self, mix_slice, Lr, Ls, chunks, last_len):
        self.Lr, self.Ls, self.C, self.last_len = Lr, Ls, chunks, last_len
        self.mix_w = restate.encoder_fwd(mix_slice.float(), self.sds["encoder"]["conv1d.weight"])   # [1, Lr, N]
        assert self.mix_w.shape[1] == Lr
        y = restate.cln_fwd(self.mix_w, self.m["layer_norm.gamma"], self.m["layer_norm.beta"])
        self.h = y @ self.m["bottleneck_conv1x1.conv.weight"][:, :, 0].t()
        self.residual = None
        self.sep_full = torch.zeros(Lr + 1, self.hp.n_spk * self.hp.enc_dim)

    def _p(self, i):
        return f"mamba_net.layers.{i}."

    def pre(self, i):
        p = self._p(i)
        self.residual = self.h if self.residual is None else self.h + self.residual
        hn = restate.rmsnorm_fwd(self.residual, self.m[p + "norm.weight"])
        xz = hn @ self.m[p + "mixer.in_proj.weight"].t()
        self.xs, self.z = xz[..., :self.di].contiguous(), xz[..., self.di:].contiguous()

    def xs_edges(self):
        return torch.stack([self.xs[0, :3], self.xs[0, self.Lr - 3:]]).contiguous()

    def conv_xproj(self, i, halo_lo, halo_hi):
        p = self._p(i) + "mixer."
        di = self.di
        lo = halo_lo.reshape(1, 3, di) if halo_lo is not None else torch.zeros(1, 3, di)
        hi = halo_hi.reshape(1, 3, di) if halo_hi is not None else torch.zeros(1, 3, di)
        xp = torch.cat([lo, self.xs, hi], dim=1)
        self.dirs = []
        for sfx, rev in (("", False), ("_b", True)):
            u = restate.causal_conv_silu(xp, self.m[p + f"conv1d{sfx}.weight"], self.m[p + f"conv1d{sfx}.bias"],
                                         reverse=rev)[:, 3:3 + self.Lr].contiguous()
            W_dt = self.m[p + f"dt_proj{sfx}.weight"]
            R = W_dt.shape[1]
            dbl = u @ self.m[p + f"x_proj{sfx}.weight"].t()
            self.dirs.append(dict(
                u=u, delta_pre=(dbl[..., :R] @ W_dt.t()).contiguous(), B=dbl[..., R:R + 16].contiguous(),
                C=dbl[..., R + 16:].contiguous(), bias=self.m[p + f"dt_proj{sfx}.bias"],
                A=-torch.exp(self.m[p + ("A_b_log" if rev else "A_log")].float()),
                D=self.m[p + ("D_b" if rev else "D")].float(), rev=rev))

    def _rows(self, c):
        return c * self.Ls, min((c + 1) * self.Ls, self.Lr)

    def _scan_chunk(self, d, c, h_in):
        a, b = self._rows(c)
        sl = lambda t: t[:, a:b].contiguous()
        return restate.selective_scan(sl(d["u"]), sl(d["delta_pre"]), d["A"], sl(d["B"]), sl(d["C"]), d["D"],
                                      sl(self.z), d["bias"], reverse=d["rev"], h_in=h_in, impl="c")

    def scan_summary(self, i):
        h_end = torch.zeros(2, self.C, self.di, 16)
        sdl = torch.zeros(2, self.C, self.di)
        for k, d in enumerate(self.dirs):
            for c in range(self.C):
                a, b = self._rows(c)
                assert b - a == (self.last_len if c == self.C - 1 else self.Ls)
                _, hl = self._scan_chunk(d, c, None)
                h_end[k, c] = hl[0]
                sdl[k, c] = F.softplus(d["delta_pre"][0, a:b] + d["bias"]).sum(0)
        return h_end, sdl

    def fold(self, i, h_end, sdl, g0, n_out, h0=None, want_final=False, dir_mask=3):
        G = h_end.shape[1]
        h_in = torch.zeros(2, n_out, self.di, 16)
        h_final = torch.zeros(2, self.di, 16) if want_final else None
        for k, d in enumerate(self.dirs):
            if not (dir_mask >> k) & 1:
                continue
            h = h0[k].clone() if h0 is not None else torch.zeros(self.di, 16)
            order = range(G) if k == 0 else range(G - 1, -1, -1)
            for g in order:
                if g0 <= g < g0 + n_out:
                    h_in[k, g - g0] = h
                h = torch.exp(d["A"] * sdl[k, g].unsqueeze(-1)) * h + h_end[k, g]
            if want_final:
                h_final[k] = h
        return h_in, h_final

    def scan_seeded(self, i, h_in):
        self.ys = []
        for k, d in enumerate(self.dirs):
            outs = [self._scan_chunk(d, c, h_in[k, c].unsqueeze(0))[0] for c in range(self.C)]
            self.ys.append(torch.cat(outs, dim=1))

    def out_proj(self, i):
        self.h = (0.5 * self.ys[0] + 0.5 * self.ys[1]) @ self.m[self._p(i) + "mixer.out_proj.weight"].t()

    def head(self):
        hp = self.hp
        self.residual = self.h + self.residual
        o = restate.rmsnorm_fwd(self.residual, self.m["mamba_net.norm_f.weight"])
        score = o @ self.m["mask_conv1x1.conv.weight"][:, :, 0].t()
        self.sep_full[1:] = (F.relu(score) * torch.cat([self.mix_w] * hp.n_spk, dim=-1))[0]

    def sep_last_row(self):
        return self.sep_full[self.Lr].clone()

    def set_sep_halo(self, row):
        self.sep_full[0] = 0 if row is None else row

    def decode(self):
        hp, N = self.hp, self.hp.enc_dim
        w = self.sds["decoder"]["weight"]
        return torch.stack([restate.decoder_fwd(self.sep_full[None, :, s * N:(s + 1) * N], w)[0]
                            for s in range(hp.n_spk)], dim=-1)
