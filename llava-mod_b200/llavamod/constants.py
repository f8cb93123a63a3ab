"""Constants of the hot path (reference: llavamod/constants.py:6-13)."""
IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_IMAGE_PATCH_TOKEN = "<im_patch>"
DEFAULT_IM_START_TOKEN = "<im_start>"
DEFAULT_IM_END_TOKEN = "<im_end>"
# hard-coded vocabulary slice of the mimic loss (reference: llavamod/train/align_trainer.py:473,497)
KD_VOCAB_SIZE = 151936
# data side (reference: llavamod/constants.py:14-22): a <video> stands for num_frames <image> tokens; per-sample caps
DEFAULT_VIDEO_TOKEN = "<video>"
DEFAULT_VID_START_TOKEN = "<vid_start>"
DEFAULT_VID_END_TOKEN = "<vid_end>"
MAX_IMAGE_LENGTH = 16
MAX_VIDEO_LENGTH = 1
