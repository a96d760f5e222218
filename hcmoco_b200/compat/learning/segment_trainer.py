"""Shadow of pycontrast/learning/segment_trainer.py (the `train_soft_joint_pri3d` path of main_segmentor.py)."""
from hcmoco_b200.segment import SegTrainer  # noqa: F401
