"""Drop-in for the reference package `e2_tts_pytorch` (src/e2_tts_pytorch/ in the reference tree), CFM sampling path only.

    from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS, DurationPredictor, MelSpec, EncodecWrapper

as src/inference_v2a.py:35-36, src/inference_v2p.py, app.py:48-49 and predict.py:48-49 do.  The arithmetic runs in
libe2b.so (hand-written sm_100a CUDA behind the C-ABI of include/e2b.h); there is no CPU or eager fallback.
"""
