"""Import the UNMODIFIED reference package from /root/reference (build container only).

TEST INFRASTRUCTURE — never imported by the product path (``ark_b200/``, ``kgvae/``).

The reference's ``kgvae.model.utils`` imports ``intelligraphs`` at module top
(/root/reference/kgvae/model/utils.py:9,11) and that package is not installable here
(no network).  We register empty stand-in modules carrying only the names the reference
imports, then put /root/reference first on ``sys.path`` so ``import kgvae`` resolves to the
reference, not to this repository's drop-in package of the same name.

Because both trees define a top-level package called ``kgvae`` this loader must run in a
process that has NOT imported this repository's ``kgvae`` — ``oracle/make_golden.py``
is that process.  It only exists in the build container: /root/reference is absent on the
GPU box, which is why the outputs are committed as fixtures under ``tests/golden/``.
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install_intelligraphs_stub():
    class _Missing:  # any attempt to actually use intelligraphs must fail loudly
        def __init__(self, *a, **k):
            raise RuntimeError("intelligraphs is stubbed: not available offline")

    root = _stub("intelligraphs", DataLoader=_Missing)
    root.evaluators = _stub("intelligraphs.evaluators", post_process_data=_Missing,
                            SemanticEvaluator=_Missing)
    root.data_loaders = _stub("intelligraphs.data_loaders", DatasetDownloader=_Missing,
                              load_data_as_list=_Missing, get_file_paths=_Missing,
                              parse_files_to_subgraphs=_Missing)
    root.verifier = _stub("intelligraphs.verifier")
    # verification.py instantiates every verifier inside get_verifier(); plain no-arg classes.
    names_syn = {n: type(n, (), {}) for n in ("SynPathsVerifier", "SynTIPRVerifier", "SynTypesVerifier")}
    names_wd = {n: type(n, (), {}) for n in ("WDMoviesVerifier", "WDArticlesVerifier")}
    root.verifier.synthetic = _stub("intelligraphs.verifier.synthetic", **names_syn)
    root.verifier.wikidata = _stub("intelligraphs.verifier.wikidata", **names_wd)


def load_reference():
    """Returns (models_module, utils_module) of the reference."""
    if "kgvae" in sys.modules and not getattr(sys.modules["kgvae"], "__file__", "").startswith(REFERENCE_ROOT):
        raise RuntimeError("this repository's kgvae is already imported; run in a fresh process")
    install_intelligraphs_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import kgvae.model.models as ref_models
    import kgvae.model.utils as ref_utils
    assert ref_models.__file__.startswith(REFERENCE_ROOT), ref_models.__file__
    return ref_models, ref_utils
