"""Token constants of the path (handsonvlm/constants.py:12-13,20; llava/constants.py:7-8)."""
IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
HAND_TOKEN_TEMPLATE = "<hand_traj>"
# 32000 Vicuna ids + 100 <t*> time tokens => <hand_traj> = 32100 (handsonvlm/train/train.py:365-368,
# hard-coded at handsonvlm/model/language_model/handsonvlm.py:146,349,609)
HAND_TRAJ_TOKEN_ID = 32100
VISUAL_TOKENS_PER_CLIP = 356          # 100 fast + 4*64 slow (handsonvlm.py:113)
