"""wfl_asr_b200 -- B200-native (sm_100a) implementation of WFL-ASR's batched labeling forward path.

Public surface mirrors the reference's files for this path: ``model.BIOPhonemeTagger``,
``infer.infer_audio`` / ``infer_folder`` / CLI, ``utils.decode_bio_tags`` /
``merge_adjacent_segments`` / ``save_lab``.  All compute runs in libwfl_b200.so (include/wfl_b200.h).
"""
__version__ = "0.1.0"
