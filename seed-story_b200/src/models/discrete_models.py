"""Drop-in for the one class of the reference's src/models/discrete_models.py that inference uses
(DiscreteModleIdentity, reference :120-130; hydra target configs/discrete_model/discrete_identity.yaml)."""
from torch import nn


class DiscreteModleIdentity(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.model = nn.Identity()

    def forward(self, image_embeds, input_ids=None, text_attention_mask=None, text_embeds=None):
        return

    def encode_image_embeds(self, image_embeds):
        return image_embeds
