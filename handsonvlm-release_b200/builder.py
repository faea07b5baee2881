"""``build_vision_tower`` -- same contract as llava/model/multimodal_encoder/builder.py:5-14: picks the tower name from
``mm_vision_tower`` / ``vision_tower`` of the config, maps bare ``clip*`` names to ``openai/...`` and constructs the
drop-in :class:`CLIPVisionTower` (so ``LlavaMetaModel.__init__`` / ``initialize_vision_modules``,
llava/model/llava_arch.py:30-32,46, work unchanged)."""
import os

from .tower import CLIPVisionTower


def build_vision_tower(vision_tower_cfg, **kwargs):
    vision_tower = getattr(vision_tower_cfg, "mm_vision_tower", getattr(vision_tower_cfg, "vision_tower", None))
    is_absolute_path_exists = os.path.exists(vision_tower)
    if not is_absolute_path_exists and os.path.basename(vision_tower).startswith("clip"):
        vision_tower = os.path.join("openai", os.path.basename(vision_tower))
    if is_absolute_path_exists or vision_tower.startswith("openai") or vision_tower.startswith("laion"):
        return CLIPVisionTower(vision_tower, args=vision_tower_cfg, **kwargs)
    raise ValueError(f"Unknown vision tower: {vision_tower}")
