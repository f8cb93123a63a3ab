"""Final-save helper (reference: llavamod/train/train_utils.py:81-117)."""
import os

import torch


def safe_save_model_for_hf_trainer(trainer, output_dir: str):
    """Adaptor pre-training writes only config.json + mm_projector.bin (under `<parent>/mm_projector/checkpoint-N.bin` when called on a
    checkpoint folder); every other stage writes the full state dict through the trainer (train_utils.py:84-117)."""
    args, model = trainer.args, trainer.model
    if getattr(args, "tune_mm_mlp_adapter", False):
        keys = ["mm_projector"] + (["embed_tokens", "embed_in"] if getattr(args, "use_im_start_end", False) else [])
        weights = {k: v.detach().cpu() for k, v in model.state_dict().items() if any(m in k for m in keys)}
        if trainer.rank == 0:
            os.makedirs(output_dir, exist_ok=True)
            model.config.save_pretrained(output_dir)
            leaf = os.path.basename(output_dir.rstrip("/"))
            if leaf.startswith("checkpoint-"):
                folder = os.path.join(os.path.dirname(output_dir.rstrip("/")), "mm_projector")
                os.makedirs(folder, exist_ok=True)
                torch.save(weights, os.path.join(folder, leaf + ".bin"))
            else:
                torch.save(weights, os.path.join(output_dir, "mm_projector.bin"))
        return
    if trainer.rank == 0:
        trainer._save(output_dir, state_dict={k: v.detach().cpu() for k, v in model.state_dict().items()})
