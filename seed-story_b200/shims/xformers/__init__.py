"""`xformers` is imported by the reference's modeling_llama_xformer.py only; the drop-in's own
src/models_clm/modeling_llama_xformer.py does not need it.  This stub exists so third-party code importing the
name does not fail; it has no attention implementation (the CUDA FMHA lives in libseedstory_b200.so)."""
