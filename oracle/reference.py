"""TEST INFRASTRUCTURE — loader for the UNMODIFIED reference FCMA path.

``oracle/build_ref.sh`` compiles ``cython_blas.pyx`` and ``fcma_extension.cc`` straight from
``/root/reference`` and copies ``voxelselector.py`` / ``classifier.py`` / ``util.py`` /
``preprocessing.py`` verbatim into ``oracle/_ref`` (git-ignored build artefacts).  This module
makes them importable without mpi4py (absent from the image): ``mpi4py`` is replaced by an
in-process stub whose ``COMM_WORLD`` reports 2 ranks (``voxelselector.py:137-139`` refuses 1),
and the master/worker message loop (``voxelselector.py:176-282``) is replaced by its serial
equivalent: one ``_voxel_scoring`` call per task — the master does no compute
(``voxelselector.py:166-168``), so this is the same arithmetic as ``mpiexec -n 2``.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(_HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "brainiak", "fcma"))


def _install_mpi_stub():
    if "mpi4py" in sys.modules and not getattr(sys.modules["mpi4py"], "_fcma_stub", False):
        return  # a real mpi4py is present; leave it alone
    mpi = types.ModuleType("mpi4py")
    mpi._fcma_stub = True
    MPI = types.ModuleType("mpi4py.MPI")

    class _Comm:
        def Get_size(self):
            return 2

        def Get_rank(self):
            return 0

        def bcast(self, obj, root=0):
            return obj

    class _Status:
        def Get_source(self):
            return 1

        def Get_tag(self):
            return 0

    MPI.COMM_WORLD = _Comm()
    MPI.Status = _Status
    MPI.ANY_SOURCE = -1
    MPI.ANY_TAG = -1
    mpi.MPI = MPI
    sys.modules["mpi4py"] = mpi
    sys.modules["mpi4py.MPI"] = MPI


_mods = None


def load():
    """Return a namespace with the reference modules (voxelselector, classifier, util,
    preprocessing, cython_blas, fcma_extension).  Raises RuntimeError if oracle/_ref is missing."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(
            "oracle/_ref not built: run `bash oracle/build_ref.sh` where /root/reference exists")
    _install_mpi_stub()
    # `brainiak` must resolve to oracle/_ref/brainiak and nothing else
    for k in [k for k in sys.modules if k == "brainiak" or k.startswith("brainiak.")]:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    try:
        import brainiak.fcma.voxelselector as voxelselector
        import brainiak.fcma.classifier as classifier
        import brainiak.fcma.util as util
        import brainiak.fcma.preprocessing as preprocessing
        import brainiak.fcma.cython_blas as cython_blas
        import brainiak.fcma.fcma_extension as fcma_extension
    finally:
        sys.path.remove(REF_ROOT)
    assert os.path.realpath(voxelselector.__file__).startswith(os.path.realpath(REF_ROOT))
    _mods = types.SimpleNamespace(
        voxelselector=voxelselector, classifier=classifier, util=util,
        preprocessing=preprocessing, cython_blas=cython_blas, fcma_extension=fcma_extension,
        VoxelSelector=voxelselector.VoxelSelector, Classifier=classifier.Classifier)
    return _mods


def run_voxel_selection(vs, clf, tasks=None):
    """Serial equivalent of VoxelSelector._master/_worker (voxelselector.py:198-238, 275-282).

    ``vs`` is a reference VoxelSelector; returns the list the master would return from run()
    (sorted by accuracy, descending, stable).  ``tasks`` optionally restricts the (start, n) list.
    """
    V = vs.num_voxels
    unit = vs.voxel_unit
    if tasks is None:
        tasks = [(s, min(unit, V - s)) for s in range(0, V, unit)]
    results = []
    for task in tasks:
        results += vs._voxel_scoring(task, clf)
    results.sort(key=lambda tup: tup[1], reverse=True)
    return results


def voxel_block_stages(vs, task, clf=None):
    """Run the reference's own stages for one task and return (corr_raw, corr_norm, kernels).

    Calls, in the order of ``_voxel_scoring`` (voxelselector.py:492-503):
    ``_correlation_computation`` -> ``fcma_extension.normalization`` -> ``_prepare_for_cross_validation``.
    """
    import sklearn.svm
    m = load()
    corr = vs._correlation_computation(task)
    raw = corr.copy()
    m.fcma_extension.normalization(corr, vs.epochs_per_subj)
    norm = corr.copy()
    if clf is None:
        clf = sklearn.svm.SVC(kernel="precomputed", shrinking=False, C=1)
    kern = vs._prepare_for_cross_validation(corr, clf)
    return raw, norm, kern
