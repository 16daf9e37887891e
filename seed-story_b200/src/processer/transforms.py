"""Drop-in for src/processer/transforms.py (reference :4-47): CPU/PIL preprocessing, not a kernel target
(SURVEY.md §8a row a1).  Only the transform types the reference defines are accepted."""
from torchvision import transforms as T

_STATS = {
    "clip": ((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711), None),
    "clipa": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225), None),
    "sd": ([0.5], [0.5], T.InterpolationMode.BICUBIC),
}


def get_transform(type='clip', keep_ratio=True, image_size=224):
    if type not in _STATS:
        raise NotImplementedError
    mean, std, interp = _STATS[type]
    kw = {} if interp is None else {"interpolation": interp}
    steps = [T.Resize(image_size, **kw), T.CenterCrop(image_size)] if keep_ratio else \
        [T.Resize((image_size, image_size), **kw)]
    return T.Compose(steps + [T.ToTensor(), T.Normalize(mean=mean, std=std)])
