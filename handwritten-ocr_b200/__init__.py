"""handwritten-ocr_b200: B200-native OCR read path (preprocess -> VLM read -> agreement/merge/CER)
behind the reference's `ocr_agent.tools` surface.  See DESIGN.md."""
__version__ = "0.1.0"


def install():
    """Route the reference's `ocr_agent.tools` hot functions to this package (see dropin.py).  The submodule is not
    called `install`: importing a submodule binds its name on the package and would replace this function."""
    from .dropin import install as _install
    return _install()


def __getattr__(name):
    # lazy: importing the package must not require CUDA or the shared library
    if name in ("preprocess_image", "run_ocr", "unload_ocr_model", "compare_versions", "merge_versions",
                "evaluate", "tier1_metrics", "cer", "wer", "levenshtein", "normalize_text", "transcribe"):
        from . import tools
        return getattr(tools, name)
    raise AttributeError(name)
