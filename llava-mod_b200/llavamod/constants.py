"""Constants of the hot path (reference: llavamod/constants.py:6-13)."""
IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_IMAGE_PATCH_TOKEN = "<im_patch>"
DEFAULT_IM_START_TOKEN = "<im_start>"
DEFAULT_IM_END_TOKEN = "<im_end>"
# hard-coded vocabulary slice of the mimic loss (reference: llavamod/train/align_trainer.py:473,497)
KD_VOCAB_SIZE = 151936
