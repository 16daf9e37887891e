"""Drop-in for src/models_clm/generation.py (reference :9-31).

The class keeps the reference's constructor and `img_ids_list`; inside seedstory_b200 its effect is applied on
the device, fused with the greedy argmax (ss_logits_process_argmax_f16), so `__call__` is only used when
someone drives it by hand on host tensors."""
import torch

BOI_TOKEN = '<img>'
EOI_TOKEN = '</img>'
IMG_TOKEN = '<img_{:05d}>'


class AutoImageTokenGenerationProcessor:
    def __init__(self, tokenizer, num_img_gen_tokens=64) -> None:
        run = BOI_TOKEN + ''.join(IMG_TOKEN.format(i) for i in range(num_img_gen_tokens)) + EOI_TOKEN
        self.img_ids_list = tokenizer.encode(run, add_special_tokens=False)

    def __call__(self, input_ids, scores):
        ids = self.img_ids_list
        for i in range(input_ids.shape[0]):
            last = int(input_ids[i, -1])
            if last in ids[:-1]:
                scores[i, ..., ids[ids.index(last) + 1]] = scores[i, ...].max() + 10.
            else:
                scores[i, ..., torch.tensor(ids[1:], dtype=torch.long)] = 0.0
        return scores


class ForcedScheduleProcessor:
    """Forces generated token k to schedule[k] when schedule[k] >= 0 (synthetic-weights benchmark schedule —
    SURVEY.md §8d; passed through the same `logits_processor=` argument the reference exposes)."""

    def __init__(self, schedule):
        self.schedule = list(schedule)


class SuppressTokensProcessor:
    """transformers' SuppressTokensLogitsProcessor (`scores[:, suppress_tokens] = -inf` at every step), accepted in
    the same `logits_processor=` list and applied on the device inside the argmax kernel.  The synthetic-weights
    benchmark uses it to keep EOS and <img> out of the free text slots of the forced turn schedule (with random
    weights any id can win a greedy step; the reference's trained model emits them where the story calls for them)."""

    def __init__(self, suppress_tokens):
        self.suppress_tokens = [int(t) for t in suppress_tokens]

    def __call__(self, input_ids, scores):
        scores[..., torch.tensor(self.suppress_tokens, dtype=torch.long)] = -float("inf")
        return scores
