"""CPU-side tests of the host layer: C-ABI library loads and exports every declared symbol, the mirror
objects behave like the reference's, and the product refuses to run without CUDA (no fallback)."""
import os
import re

import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from voicefixer_main_b200 import _lib
    lib = _lib.load_library()
    header = open(os.path.join(ROOT, "include", "b200vf.h")).read()
    declared = re.findall(r"VF_API\s+[\w\s\*]+?\b(vf_\w+)\s*\(", header)
    assert len(declared) >= 20
    assert sorted(declared) == sorted(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    from voicefixer_main_b200 import VoiceFixer
    from voicefixer_main_b200._lib import EngineError
    m = VoiceFixer()
    with pytest.raises(RuntimeError):
        m.to("cpu")
    with pytest.raises(RuntimeError):
        m.restore(torch.zeros(1, 4410))
    if not torch.cuda.is_available():
        with pytest.raises((EngineError, RuntimeError)):
            m.to("cuda:0")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "voicefixer_main_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_hparams_and_filterbank_mirror_reference_semantics(tmp_path):
    from voicefixer_main_b200.model import HParams, get_hparams_from_file, melscale_fbanks
    from oracle import vf_oracle as O
    p = tmp_path / "c.json"
    p.write_text('{"model": {"window_size": 2048, "hop_size": 441}, "data": {"sampling_rate": 44100}}')
    hp = get_hparams_from_file(str(p))
    assert hp["model"]["hop_size"] == 441 and hp.model.window_size == 2048 and "data" in hp and len(hp) == 2
    assert isinstance(hp.model, HParams)
    assert torch.equal(melscale_fbanks(), O.mel_filterbank())


def test_unsupported_configs_fail_loudly():
    from voicefixer_main_b200 import VoiceFixer, default_hparams
    hp = default_hparams()
    hp["task"]["gsr"]["gsr_model"]["voicefixer"]["unet"] = False
    with pytest.raises(NotImplementedError):
        VoiceFixer(hp)
    hp = default_hparams()
    hp["model"]["window_size"] = 1024
    with pytest.raises(NotImplementedError):
        VoiceFixer(hp)


def test_arch_keys_match_reference_state_dict():
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    from voicefixer_main_b200.arch import UNET_PREFIX, unet_keys
    from voicefixer_main_b200.weights import make_state
    model, _ = ref_import.build_reference_model(make_state(1234))
    own = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith(UNET_PREFIX)}
    mine = {UNET_PREFIX + k: tuple(s) for k, s in unet_keys()}
    assert own == mine
    assert list(own) == list(mine)          # same registration order
    assert sum(int(torch.tensor(s).prod()) if s else 1 for k, s in mine.items()
               if not k.endswith("num_batches_tracked") and "running" not in k) == 65152867
