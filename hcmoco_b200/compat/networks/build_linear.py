"""Shadow of pycontrast/networks/build_linear.py (`build_segmentor`: the FCN head of main_segmentor.py:40)."""
from hcmoco_b200.segment import FCNHead, build_segmentor  # noqa: F401
