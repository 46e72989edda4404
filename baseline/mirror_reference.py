"""Mirrors the handful of UNMODIFIED reference files that the CPU reference arm of bench.py executes into the
git-ignored ``baseline/_ref/`` (it travels to the GPU box with the snapshot; /root/reference does not exist there).

Run by ``__graft_entry__.build()`` in the build container.  Nothing here is product code and nothing is edited: the
files are byte copies, listed below with the entry point each one provides.  The reference is not pip-installable (no
setup.py / pyproject) and its Sinkhorn dependency ``geomloss==0.2.4`` (requirements.txt:1) cannot be installed offline,
so the arm imports ``oracle.geomloss_ref`` under the name ``geomloss`` -- hence ``kind: "reference+geomloss-stub"``.
"""
import os
import shutil

REFERENCE_ROOT = os.environ.get("ASPIRE_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")

FILES = [
    # the release API the evaluation code imports (src/evaluation/utils/models.py:2-3); CPU by construction: the
    # .cuda() lines of the training twin are commented out here (examples/ex_aspire_consent_multimatch.py:141-142)
    "examples/__init__.py",
    "examples/ex_aspire_consent_multimatch.py",   # AllPairMaskedWasserstein.compute_distance (:118-189)
    "examples/ex_aspire_consent.py",              # AspireConSent / prepare_abstracts (:25-212)
    # the training-side twin that caching_score calls (disent_models.py:294-297)
    "src/__init__.py",
    "src/learning/__init__.py",
    "src/learning/facetid_models/__init__.py",
    "src/learning/facetid_models/pair_distances.py",   # compute_distance (:21-92), allpair_masked_dist_l2max (:138-186)
    "src/learning/models_common/__init__.py",
    "src/learning/models_common/activations.py",
]


def mirror():
    """Copy FILES from the reference tree; returns DEST, or None when the reference tree is absent (GPU box)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "learning")):
        return DEST if os.path.isdir(DEST) else None
    for rel in FILES:
        src, dst = os.path.join(REFERENCE_ROOT, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    with open(os.path.join(DEST, "MIRROR.txt"), "w") as fh:
        fh.write("byte copies of allenai/aspire files made by baseline/mirror_reference.py; not tracked by git\n")
        fh.write("\n".join(FILES) + "\n")
    return DEST


if __name__ == "__main__":
    print(mirror())
