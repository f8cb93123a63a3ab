"""Prompt templates of the distillation recipes (reference: llavamod/conversation.py).

Only what `--version qwen` / `phi` / `stablelm` (two-separator chat, conversation.py:319-329,461) and `--version plain`
(adaptor pre-training, :393-402,467-468) need: a rendered prompt string whose token boundaries the label masking of
data/data_utils.py relies on.  `default_conversation` is module state set by the train entry points from `--version`, exactly as the
reference's entry points do (align_train.py:438-441)."""
import dataclasses
from enum import Enum, auto
from typing import List, Sequence, Tuple


class SeparatorStyle(Enum):
    TWO = auto()       # "<system><sep>USER: q<sep>ASSISTANT: a<sep2>USER: ..."
    PLAIN = auto()     # "<image>caption<sep>"


@dataclasses.dataclass
class Conversation:
    system: str
    roles: Tuple[str, str]
    messages: List[Sequence[str]]
    offset: int = 0
    sep_style: SeparatorStyle = SeparatorStyle.TWO
    sep: str = " "
    sep2: str = None
    version: str = "unknown"

    def append_message(self, role, message):
        self.messages.append([role, message])

    def copy(self):
        return dataclasses.replace(self, messages=[list(m) for m in self.messages])

    def get_prompt(self):
        """conversation.py:52-62 (TWO) and :93-99 (PLAIN).  A message that is None renders as the bare "ROLE:" generation prompt."""
        if self.sep_style is SeparatorStyle.TWO:
            closers = (self.sep, self.sep2)
            out = [self.system, closers[0]]
            for turn, (role, text) in enumerate(self.messages):
                out.append("%s: %s%s" % (role, text, closers[turn % 2]) if text else role + ":")
            return "".join(out)
        if self.sep_style is SeparatorStyle.PLAIN:
            closers = (self.sep, self.sep2 if self.sep2 is not None else self.sep)
            return self.system + "".join((text + closers[turn % 2]) if text else "" for turn, (_, text) in enumerate(self.messages))
        raise ValueError("unsupported separator style %r" % (self.sep_style,))


_CHAT_SYSTEM = ("A chat between a curious user and an artificial intelligence assistant. "
                "The assistant gives helpful, detailed, and polite answers to the user's questions.")

conv_phi = Conversation(system=_CHAT_SYSTEM, roles=("USER", "ASSISTANT"), messages=[], version="phi",
                        sep_style=SeparatorStyle.TWO, sep=" ", sep2="<|endoftext|>")
conv_stablelm = dataclasses.replace(conv_phi, version="stablelm")
conv_llava_plain = Conversation(system="", roles=("", ""), messages=[], version="plain", sep_style=SeparatorStyle.PLAIN, sep="\n")

# `--version qwen` maps to the phi template (reference conversation.py:461)
conv_templates = {"phi": conv_phi, "qwen": conv_phi, "stablelm": conv_stablelm, "plain": conv_llava_plain, "v0_plain": conv_llava_plain}
default_conversation = conv_phi


def set_default_conversation(version):
    """Entry points call this with --version.  Templates of other model families (vicuna, llama-2, mpt, gemma-2, openchat...) belong
    to recipes outside the Qwen distillation path and are not carried."""
    global default_conversation
    if version not in conv_templates:
        raise NotImplementedError("conversation template %r is not part of the Qwen distillation path (have: %s)" % (version, sorted(conv_templates)))
    default_conversation = conv_templates[version]
    return default_conversation
