"""The Python examples of INTEGRATION.md section 4 run as written (at reduced sizes): documentation that cannot rot."""
import os
import re

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_md_quick_start_blocks_run():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    section = text[text.index("## 4. Python quick start"):]
    section = section[:section.index("\n## ", 5)] if "\n## " in section[5:] else section
    blocks = re.findall(r"```python\n(.*?)```", section, flags=re.S)
    assert len(blocks) >= 4
    ns = {}
    for code in blocks:                                  # the blocks share one namespace, top to bottom, as a reader would
        code = code.replace("1 << 20", "4096").replace("65536", "2048").replace("T=128", "T=16")
        if "actions_cpu_int8" in code and "actions_cpu_int8" not in ns:
            ns["actions_cpu_int8"] = torch.randint(0, 4, (8, 4096), dtype=torch.int8)
        exec(compile(code, "INTEGRATION.md", "exec"), ns)
    torch.cuda.synchronize()
    assert ns["batch"]["obs"].shape == (17, 2048, 8)
    assert ns["ev"]["reward"].shape == (8, 4096) and ns["ring8"].appended == 16 and ns["ring"].appended == 16
    assert ns["res"].reward.shape[0] == 8
