"""Stage the UNMODIFIED reference (rhuangr/MARL-Maze) under baseline/_ref/ so that it travels to the GPU box.

    python baseline/stage_reference.py            # copies /root/reference/{*.py,PPO.pth} -> baseline/_ref/

baseline/_ref/ is git-ignored (reference sources are never committed) but NOT gpurun-ignored, so `gpurun` ships it with the
snapshot.  The reference has no setup.py / pyproject.toml -- `pip install --target baseline/_ref /root/reference` has nothing to
build -- so staging is a plain file copy of its five modules and its checkpoint.  __graft_entry__.build() calls this when
/root/reference is present (the build container); on the GPU box the staged copy is used as it is.
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("MARL_MAZE_REFERENCE", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ["main.py", "maze.py", "maze_agent.py", "PPO.py", "networks.py", "PPO.pth"]


def staged() -> bool:
    return all(os.path.isfile(os.path.join(REF_DST, f)) for f in FILES)


def stage(force: bool = False) -> str | None:
    """Returns the staged directory, or None when neither the reference nor a staged copy exists."""
    if not os.path.isfile(os.path.join(REF_SRC, "maze.py")):
        return REF_DST if staged() else None
    os.makedirs(REF_DST, exist_ok=True)
    for f in FILES:
        src, dst = os.path.join(REF_SRC, f), os.path.join(REF_DST, f)
        if force or not os.path.isfile(dst) or os.path.getmtime(dst) < os.path.getmtime(src) or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copy2(src, dst)
    return REF_DST


if __name__ == "__main__":
    print(stage(force=True))
