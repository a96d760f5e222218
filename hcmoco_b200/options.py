"""Launcher-compatible flags: the subset of pycontrast/options/base_options.py:31-151 and train_options.py:8-74 that the
RGBD2S pre-train scripts pass, with the same names, defaults, `--method` override table (base_options.py:12-22, 168-177)
and derived fields (model_name / model_folder / tb_folder / warm-up, train_options.py:26-74).  Unlike the reference's
parser it accepts `--method CMCJointsPri3DRGBD2S` and `--mem bank+jointspri3d`, which the shipped second-stage scripts
pass but the reference's `choices` lists reject."""
import argparse
import math
import os

OVERRIDE = {   # method: modal, jigsaw, mem, aug, head, nce_t
    "CMCRGBD2S": ["RGBD2S", False, "bank", "C", "linear", 0.07],
    "CMCJointsPri3DRGBD2S": ["RGBD2S", False, "bank+jointspri3d", "C", "linear", 0.07],
}


class TrainOptions(object):
    def initialize(self, p):
        a = p.add_argument
        a("--data_folder", type=str, default="./data"); a("--train_file_list", type=str, default="")
        a("--val_file_list", type=str, default=""); a("--model_path", type=str, default="./save")
        a("--tb_path", type=str, default="./tb"); a("--pretrain", type=str, default=None); a("--tag", type=str, default="")
        a("--print_freq", type=int, default=10); a("--save_freq", type=int, default=20)
        a("--batch_size", type=int, default=256); a("-j", "--num_workers", type=int, default=40)
        a("--epochs", type=int, default=200); a("--learning_rate", type=float, default=0.03)
        a("--lr_decay_epochs", type=str, default="120,160"); a("--lr_decay_rate", type=float, default=0.1)
        a("--weight_decay", type=float, default=1e-4); a("--momentum", type=float, default=0.9)
        a("--cosine", action="store_true"); a("--downstream_training", action="store_true")
        a("--method", default="Customize", type=str, choices=["Customize"] + list(OVERRIDE))
        a("--modal", default="RGBD2S", type=str, choices=["RGBD2S"]); a("--in_channel_list", type=str, default="3,3")
        a("--linear_feat_map", type=int, default=0); a("--width", type=int, default=18); a("--dataset", type=str, default="")
        a("--IN_Pretrain", type=str, default=None); a("--depth_Pretrain", type=str, default=None)
        a("--pri3d_num_samples_per_image", type=int, default=400); a("--modality_missing", type=int, default=0)
        a("--mpii_root", type=str, default=""); a("--coco_root", type=str, default=""); a("--pool_method", type=str, default="mean")
        a("--cmc_loss_weight", type=float, default=1.0); a("--skeleton_meta_name", type=str, default="mpii")
        a("--not_use_weighted_sampler", action="store_true", default=False); a("--temperature", type=float, default=0.07)
        a("--random_flip", type=int, default=0); a("--jigsaw", action="store_true")
        a("--mem", default="bank", type=str, choices=["bank", "bank+jointspri3d"]); a("--arch", default="HRNet", type=str)
        a("-d", "--feat_dim", default=128, type=int); a("-k", "--nce_k", default=65536, type=int)
        a("-m", "--nce_m", default=0.5, type=float); a("-t", "--nce_t", default=0.07, type=float)
        a("--alpha", default=0.999, type=float); a("--head", default="linear", type=str); a("--resume", default="", type=str)
        a("--world-size", default=-1, type=int); a("--rank", default=-1, type=int)
        a("--dist-url", default="tcp://127.0.0.1:23456", type=str); a("--dist-backend", default="nccl", type=str)
        a("--seed", default=None, type=int); a("--gpu", default=None, type=int)
        a("--multiprocessing-distributed", action="store_true")
        a("--aug", default="A", type=str); a("--beta", type=float, default=0.5); a("--warm", action="store_true")
        a("--amp", action="store_true"); a("--opt_level", type=str, default="O2"); a("--n_class", type=int, default=31)
        return p

    def parse(self, argv=None, make_dirs=True):
        p = self.initialize(argparse.ArgumentParser("hcmoco_b200 pre-train (launcher-compatible)"))
        opt = p.parse_args(argv)
        if opt.method in OVERRIDE:
            opt.modal, opt.jigsaw, opt.mem, opt.aug, opt.head, opt.nce_t = OVERRIDE[opt.method]
        if opt.amp:
            raise NotImplementedError("--amp (apex) is not on the engine's path: single-pass low precision fails the "
                                      "1e-3 parity bar (SURVEY.md F8)")
        opt.lr_decay_epochs = [int(v) for v in opt.lr_decay_epochs.split(",")]
        opt.in_channel_list = [int(v) for v in opt.in_channel_list.split(",")]
        opt.model_name = "{}_{}_{}_Jig_{}_{}_aug_{}_{}_{}_{}".format(opt.method, opt.arch, opt.modal, opt.jigsaw, opt.mem,
                                                                     opt.aug, opt.head, opt.nce_t, opt.tag)
        if opt.cosine:
            opt.model_name += "_cosine"
        if opt.batch_size > 256:
            opt.warm = True
        if opt.warm:
            opt.model_name += "_warm"
            opt.warmup_from = 0.01
            opt.warm_epochs = 10 if opt.epochs > 500 else 5
            if opt.cosine:
                eta_min = opt.learning_rate * (opt.lr_decay_rate ** 3)
                opt.warmup_to = eta_min + (opt.learning_rate - eta_min) * (1 + math.cos(math.pi * opt.warm_epochs / opt.epochs)) / 2
            else:
                opt.warmup_to = opt.learning_rate
        opt.model_folder = os.path.join(opt.model_path, opt.model_name)
        opt.tb_folder = os.path.join(opt.tb_path, opt.model_name)
        if make_dirs:
            os.makedirs(opt.model_folder, exist_ok=True)
            os.makedirs(opt.tb_folder, exist_ok=True)
        return opt
