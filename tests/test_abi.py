"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/ustrun.h declares; modules keep the reference's state_dict layout; the product path has no
CPU fallback (it fails loudly) and never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import PKG, ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ustrun.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ustrun_[a-zA-Z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(PKG, "libustrun_sm100.so"))
    syms = _declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ustrun.h but not exported"
    assert lib.ustrun_abi_version() == 1


def test_binding_covers_header():
    from ustrun import _lib
    assert sorted(_lib.EXPORTS) == _declared_symbols()


def test_library_has_blackwell_instructions():
    """tcgen05 / TMA must be in the SASS (UTCHMMA, UTMALDG, UTMASTG, LDTM)."""
    so = os.path.join(PKG, "libustrun_sm100.so")
    try:
        out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, timeout=300).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnemonic in out, mnemonic


def test_no_cpu_path():
    from networks.unet_model import UNet
    from utils.losses import DiceLossWithMask
    m = UNet(1, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 1, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU path"):
        DiceLossWithMask(2)(torch.zeros(1, 2, 8, 8), torch.zeros(1, 1, 8, 8), softmax=True)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/", "").lower() or f == "__init__.py" or "import oracle" not in text, f
                assert "from oracle" not in text and "import oracle" not in text, f"{f} imports the oracle"
    # tools/ are product-side utilities too; bench.py may use the oracle only in its CPU legs
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            text = open(os.path.join(ROOT, "tools", f)).read()
            assert "from oracle" not in text and "import oracle" not in text, f"tools/{f} imports the oracle"
    bench = open(os.path.join(ROOT, "bench.py")).read()
    gpu_arm = bench[bench.index("def run_ours("):bench.index("def oracle_models(")]      # the baseline legs (cpu / eager-gpu / reference arm) follow
    assert gpu_arm and "oracle" not in gpu_arm.replace("no oracle on this arm", ""), "bench.py GPU arm touches the oracle"


def test_state_dict_layout_matches_oracle_init():
    from networks import unet as B
    from networks.unet_model import UNet
    from oracle import unet_ref as U
    torch.manual_seed(1337)
    m = UNet(3, 3)
    st = U.init_unet_a(3, 3, seed=1337)
    sd = m.state_dict()
    assert list(sd) == list(st) and len(sd) == 118
    assert all(torch.equal(sd[k], st[k]) for k in sd)
    assert len(list(m.parameters())) == 64
    torch.manual_seed(1337)
    b = B.UNet(1, 2)
    st = U.init_unet_b(1, 2, seed=1337)
    sd = b.state_dict()
    assert list(sd) == list(st) and len(sd) == 184
    assert all(torch.equal(sd[k], st[k]) for k in sd)
    assert len(list(b.parameters())) == 106
    for mk, init in ((lambda: B.UNet(3, 2, norm="dsbn", num_domains=3), lambda: U.init_unet_b(3, 2, seed=1337, norm="dsbn", num_domains=3)),
                     (lambda: UNet(3, 2, norm="dsbn", num_domains=3), lambda: U.init_unet_a(3, 2, seed=1337, norm="dsbn", num_domains=3))):
        torch.manual_seed(1337)
        sd, st = mk().state_dict(), init()
        assert list(sd) == list(st) and all(torch.equal(sd[k], st[k]) for k in sd)
    torch.manual_seed(1337)
    r = B.Rec_Decoder(num_classes=2, norm="dsbn", num_domains=3)
    st = U.init_rec_decoder(num_classes=2, norm="dsbn", num_domains=3, seed=1337)
    sd = r.state_dict()
    assert list(sd) == list(st) and all(torch.equal(sd[k], st[k]) for k in sd)


def test_error_conventions():
    from networks import unet as B
    from networks.dsbn import DomainSpecificBatchNorm2d
    with pytest.raises(ValueError, match="not supporter"):
        B.normalization(8, "xx")
    d = DomainSpecificBatchNorm2d(8, 3)
    with pytest.raises(ValueError, match="expected 4D input"):
        d(torch.zeros(2, 8, 4), torch.tensor([0]))
    with pytest.raises(TypeError):
        B.UNet(norm="dsbn")         # upstream: num_domains=None -> TypeError at construction


def test_synthetic_workload_matches_test_generator():
    """bench.py's GPU arm and tools/ draw their inputs from ustrun.synth (the measured arm never imports the
    oracle); the cpu_baseline / reference arm draws from the oracle's generator.  Same seeds -> same workload."""
    import torch
    from oracle import ssl_step_ref as S
    from oracle import unet_ref as U
    from ustrun import synth
    for br in ("softmax", "sigmoid"):
        a = synth.synthetic_batch(3, 3, 32, 48, 2, 3, seed=11, branch=br)
        b = S.synthetic_batch(3, 3, 32, 48, 2, 3, seed=11, branch=br)
        assert a.keys() == b.keys()
        for k in a:
            assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
    assert synth.conv_flops_unet_a(1, 2, 384, 384) == U.conv_flops_unet_a(1, 2, 384, 384)
