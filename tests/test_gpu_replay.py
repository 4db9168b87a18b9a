"""Config 1 in miniature: the C++ replay harness (tools/replay_config1, built by `make` against the reference's own
libmultiexp.h when that tree is present) drives the C-ABI exactly as Porla's glue does -- GoSlices over heap buffers -- through
the update / hierarchy-rebuild / C-rebuild / audit call census of SURVEY.md Appendix C, three times: legacy per-call symbols,
batched symbols, CPU restatement.  All three must end with byte-identical MAC arrays and audit replies, and every KZG proof
must verify."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tools", "replay_config1")


@pytest.mark.parametrize("blocks", [16, 64])
def test_replay_passes_agree(blocks):
    if not os.path.exists(EXE):
        pytest.skip("tools/replay_config1 not built")
    from oracle import loader
    loader.bn254()                                   # makes sure oracle/liboracle_bn254.so exists
    p = subprocess.run([EXE, "--blocks", str(blocks), "--audits", "3", "--cpu-audits", "3", "--oracle",
                        os.path.join(ROOT, "oracle", "liboracle_bn254.so")], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    d = json.loads(p.stdout)
    assert d["legacy_equals_batched_state"] is True and d["legacy_equals_batched_audits"] is True
    assert d["cpu_equals_legacy"] is True
    for name in ("legacy", "batched", "cpu"):
        assert d[name]["proofs_verified"] is True and d[name]["updates"] == blocks
    # the hierarchy rebuilds really ran: 4 (2^L - 1) butterflies per update that merges L levels, plus the C rebuild
    assert d["legacy"]["butterflies"] == d["batched"]["butterflies"] > blocks
