"""Constants of the hot path and of the data side.  Names and values are part of the drop-in surface (reference: llavamod/constants.py)."""
# integer sentinels carried by the batch tensors: ignored label, image placeholder inside input_ids
IGNORE_INDEX, IMAGE_TOKEN_INDEX = -100, -200

# placeholder strings of the conversation records; a <video> stands for `num_frames` <image> tokens
DEFAULT_IMAGE_TOKEN, DEFAULT_VIDEO_TOKEN = "<image>", "<video>"
DEFAULT_IMAGE_PATCH_TOKEN = "<im_patch>"
DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN = "<im_start>", "<im_end>"
DEFAULT_VID_START_TOKEN, DEFAULT_VID_END_TOKEN = "<vid_start>", "<vid_end>"

# per-record media caps of preprocess_multimodal / the lazy datasets
MAX_IMAGE_LENGTH, MAX_VIDEO_LENGTH = 16, 1

# hard-coded vocabulary slice of the mimic loss (reference: llavamod/train/align_trainer.py:473,497)
KD_VOCAB_SIZE = 151936
