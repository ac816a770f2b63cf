"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports exactly the
symbols include/vgslm.h declares; the ctypes structs mirror the C layouts; state-dict contract."""
import ctypes
import json
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "vgslm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vg_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from vae_gslm_b200 import _lib
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vgslm.h but not exported by libvgslm.so"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes signature table and header disagree"
    assert lib.vg_version() == 100


def test_entry_points_fail_loudly_without_gpu_or_with_bad_arguments():
    from vae_gslm_b200 import _lib
    lib = _lib.load()
    # argument validation happens before any CUDA call: null pointers → negative rc + message
    rc = lib.vg_rmsnorm_fwd(None, None, None, None, None, 4, 64, 1e-6, 0, 0, None)
    assert rc < 0 and b"null pointer" in lib.vg_last_error_string()
    g = _lib.GemmArgs()
    assert lib.vg_gemm(ctypes.byref(g), 0, None, 0, None) < 0
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            from vae_gslm_b200 import ops
            ops.rmsnorm(torch.randn(2, 64), torch.ones(64), 1e-6)


def test_ctypes_structs_match_c_layout(tmp_path):
    """compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors"""
    from vae_gslm_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "vgslm.h"
int main(void) {
  printf("{\"gemm\": [%zu, %zu, %zu, %zu], \"front\": [%zu, %zu, %zu], \"front_bwd\": [%zu, %zu],"
         " \"back\": [%zu, %zu, %zu], \"back_bwd\": [%zu, %zu]}\n",
         sizeof(vg_gemm_args), offsetof(vg_gemm_args, bias), offsetof(vg_gemm_args, row_mask), offsetof(vg_gemm_args, beta),
         sizeof(vg_latent_front_args), offsetof(vg_latent_front_args, temperature), offsetof(vg_latent_front_args, act_dtype),
         sizeof(vg_latent_front_bwd_args), offsetof(vg_latent_front_bwd_args, d_z),
         sizeof(vg_latent_back_args), offsetof(vg_latent_back_args, ln_eps), offsetof(vg_latent_back_args, kl_sum),
         sizeof(vg_latent_back_bwd_args), offsetof(vg_latent_back_bwd_args, d_log_p));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    c = json.loads(subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout)
    G, F, FB, B, BB = _lib.GemmArgs, _lib.LatentFrontArgs, _lib.LatentFrontBwdArgs, _lib.LatentBackArgs, _lib.LatentBackBwdArgs
    assert c["gemm"] == [ctypes.sizeof(G), G.bias.offset, G.row_mask.offset, G.beta.offset]
    assert c["front"] == [ctypes.sizeof(F), F.temperature.offset, F.act_dtype.offset]
    assert c["front_bwd"] == [ctypes.sizeof(FB), FB.d_z.offset]
    assert c["back"] == [ctypes.sizeof(B), B.ln_eps.offset, B.kl_sum.offset]
    assert c["back_bwd"] == [ctypes.sizeof(BB), BB.d_log_p.offset]


def test_state_dict_contract_matches_reference_checkpoint_keys(golden):
    """same keys, same order, same shapes as the reference LVTR (fixture generated from the real reference)"""
    import copy
    from vae_gslm_b200.hparams.hp import Hparams
    from vae_gslm_b200.models.speech.lvtr import LVTR
    model = LVTR(Hparams.from_dict(copy.deepcopy(golden["config"])), input_dim=golden["n_mels"])
    mine = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    ref = {k: tuple(v.shape) for k, v in golden["state_dict"].items()}
    assert set(ref) <= set(mine)
    assert all(mine[k] == ref[k] for k in ref)
    extra = set(mine) - set(ref)
    assert all(k.startswith("decoder.") and "." not in k[len("decoder."):] for k in extra)   # schedule buffers only
    assert set(golden["grads"]) == {n for n, _ in model.named_parameters()}


def test_full_config_parameter_count():
    from vae_gslm_b200.hparams.hp import Hparams
    from vae_gslm_b200.models.speech.lvtr import LVTR
    hp = Hparams.from_yamlfile(os.path.join(ROOT, "vae_gslm_b200", "configs", "train", "speech", "vae-gslm.yaml"))
    with torch.device("meta"):
        model = LVTR(hp.model, input_dim=80)
    assert sum(p.numel() for p in model.parameters()) == 226_957_564
    sd = model.state_dict()
    assert sd["transformer.0.layers.15.self_attn.in_proj.weight"].shape == (3072, 1024)
    assert sd["transformer.0.layers.0.linear1.bias"].shape == (4096,)     # FFN biased, attention not
    assert "transformer.0.layers.0.self_attn.in_proj.bias" not in sd
    assert sd["transformer.0.final_norm.scale"].shape == (1024,)
