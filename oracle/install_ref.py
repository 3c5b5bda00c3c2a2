"""Install the reference's own hot-path files, UNMODIFIED, into ``baseline/_ref/`` so that the reference arm of
``bench.py`` can run the reference's literal model code on the GPU box (which has no ``/root/reference``).

  python oracle/install_ref.py          (run by __graft_entry__.build() when /root/reference is present)

The reference cannot be pip-installed (no setup.py / pyproject) and imports ``dgl`` /
``torch_geometric``, which are absent; ``oracle/shim.py`` supplies stand-ins for those two packages and imports
the files below by path.  ``baseline/_ref/`` is git-ignored: nothing of the reference enters the repository's history.
TEST / BASELINE INFRASTRUCTURE ONLY -- the product never imports from it.
"""
import os
import shutil
import sys

SRC = "/root/reference/immunostruct"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "immunostruct")
FILES = ["models/__init__.py", "models/mapping.py", "models/layers.py", "models/hybrid_models.py",
         "models/comparative_models.py", "models/ablation_models.py", "utils/__init__.py", "utils/loss.py",
         "utils/contrastive.py", "utils/scheduler.py", "utils/seed.py", "utils/update_paths.py"]


def install() -> str:
    if not os.path.isdir(SRC):
        return ""
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            if rel.endswith("__init__.py"):                 # namespace package in the reference: keep it importable
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                open(dst, "a").close()
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    with open(os.path.join(os.path.dirname(DST), "README"), "w") as f:
        f.write("Unmodified copies of the reference's hot-path files, written by oracle/install_ref.py; git-ignored.\n")
    return DST


if __name__ == "__main__":
    print(install() or "no /root/reference on this box: nothing installed", file=sys.stderr)
