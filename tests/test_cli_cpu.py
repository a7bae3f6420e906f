"""The driver shim accepts the reference's own command lines (scripts/target/*.bash -> main_target.py:28-82)."""
import glob
import os
import shlex

import pytest

from vae_segmentation_b200 import main_target as cli

REF_SCRIPTS = "/root/reference/scripts/target"
# argument vector of scripts/target/domain_msd_dh_ft1.bash (BASELINE.json configs[3]); kept here for boxes without the tree
DH_FT1 = ("domain_msd_dh_ft1 -G 0 --method domain_adaptation --load_prefix seg_nih --load_prefix_vae vae_nih --train_list MSD_train "
          "--val_list MSD_val --data_root X --val_data_root X --data_path data/Multi_all.json --pan_index 10 --lambda_vae 1.0 "
          "--domain_loss_type 8 --val_finetune 1 --eval_epoch 2 --save_epoch 100 --max_epoch 50")


def test_parser_accepts_the_shipped_preset():
    a = cli.build_parser().parse_args(shlex.split(DH_FT1))
    assert a.method == "domain_adaptation" and a.domain_loss_type == 8 and a.val_finetune == 1 and a.lambda_vae == 1.0
    assert cli.mask_index_from(a.pan_index) == [[0, 0], [[1, 2], 1]]
    assert cli.mask_index_from("1,3") == [[0, 0], [1, 1], [3, 2]]


@pytest.mark.skipif(not os.path.isdir(REF_SCRIPTS), reason="reference tree not present")
def test_parser_accepts_every_reference_target_script():
    scripts = sorted(glob.glob(os.path.join(REF_SCRIPTS, "*.bash")))
    assert len(scripts) >= 8
    for path in scripts:
        text = open(path).read().replace("\\\n", " ")
        line = [l for l in text.splitlines() if "main_target.py" in l][0]
        argv = shlex.split(line.replace("$1", "0").replace("<Your_MSD_data_path>", "X").replace("<Your_Synapse_data_path>", "X"))
        argv = [t for t in argv[argv.index("main_target.py") + 1:]]
        argv = [("X" if (t.startswith("<") and t.endswith(">")) else t) for t in argv]
        a = cli.build_parser().parse_args(argv)
        assert a.method == "domain_adaptation", path


def test_no_cuda_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(SystemExit, match="no CUDA device"):
        cli.main(shlex.split(DH_FT1) + ["--synthetic", "2"])


# ---- main_source.py (scripts/source/*.bash) ---------------------------------------------------------------------------
REF_SOURCE_SCRIPTS = "/root/reference/scripts/source"
SEG_NIH = ("seg_nih -G 0 --method seg_train --train_list NIH_train --val_list NIH_val --data_root X --val_data_root X "
           "--data_path data/Multi_all.json --eval_epoch 20 --save_epoch 800 --max_epoch 2400")


def test_source_parser_accepts_the_shipped_preset():
    from vae_segmentation_b200 import main_source as src
    a = src.build_parser().parse_args(shlex.split(SEG_NIH))
    assert a.method == "seg_train" and a.eval_epoch == 20 and a.save_epoch == 800 and a.max_epoch == 2400 and a.lr_seg == 1e-2
    d = src.build_parser().parse_args(["p"])
    assert d.method == "vae_train" and d.batch_size == 4 and d.lr_vae == 0          # main_source.py:29,35,49 defaults


@pytest.mark.skipif(not os.path.isdir(REF_SOURCE_SCRIPTS), reason="reference tree not present")
def test_source_parser_accepts_every_reference_source_script():
    from vae_segmentation_b200 import main_source as src
    scripts = sorted(glob.glob(os.path.join(REF_SOURCE_SCRIPTS, "*.bash")))
    assert len(scripts) == 2
    methods = set()
    for path in scripts:
        text = open(path).read().replace("\\\n", " ").rstrip("\\ \n")          # the scripts end in a dangling backslash
        line = [l for l in text.splitlines() if "main_source.py" in l][0]
        argv = shlex.split(line.replace("$1", "0"))
        argv = [("X" if (t.startswith("<") and t.endswith(">")) else t) for t in argv[argv.index("main_source.py") + 1:]]
        methods.add(src.build_parser().parse_args(argv).method)
    assert methods == {"seg_train", "vae_train"}


def test_source_unwired_method_and_missing_gpu_are_errors():
    import torch
    from vae_segmentation_b200 import main_source as src
    with pytest.raises(SystemExit, match="not wired"):
        src.main(["p", "--method", "joint_train", "--synthetic", "2"])
    if not torch.cuda.is_available():
        with pytest.raises(SystemExit, match="no CUDA device"):
            src.main(shlex.split(SEG_NIH) + ["--synthetic", "2"])


# ---- bench.py accounting ------------------------------------------------------------------------------------------------
def test_bench_kernel_cost_covers_the_hot_entry_points():
    """bench.py's roofline object is built from kernel_cost(entry point, integer-argument key): every entry point that can
    be the dominant kernel of a mode must have an algorithmic (flops, bytes) model -- a zero here would silently report
    frac = 0.  Keys as the C-ABI calls produce them (integer arguments in call order)."""
    import bench
    vox = 2 * 96 ** 3
    fl, by = bench.kernel_cost("vs_conv3x3x3_tc_kdn_ex", (1, 2, 96, 96, 96, 8, 8))
    assert fl == 2.0 * 27 * 8 * 8 * vox and by == vox * 16 * 2
    fl, by = bench.kernel_cost("vs_conv3x3x3_tc_kdn_planar", (1, 2, 96, 96, 96, 8))
    assert fl == 2.0 * 27 * 8 * 8 * vox and by == vox * (16 + 8)
    fl, by = bench.kernel_cost("vs_conv3x3x3_fprop", (1, 1, 0, 0, 1, 2, 24, 24, 24, 32, 32))
    assert fl == 2.0 * 27 * 32 * 32 * 2 * 24 ** 3 and by == 2 * 24 ** 3 * 64 * 2
    fl, by = bench.kernel_cost("vs_conv3x3x3_wgrad", (1, 0, 0, 1, 2, 48, 48, 48, 16, 16))
    assert fl > 0 and by == 2 * 48 ** 3 * 32 * 2
    for name in ("vs_k2s2_gather_tc", "vs_k2s2_scatter_tc"):
        fl, by = bench.kernel_cost(name, (2, 48, 48, 48, 16, 16))
        assert fl == 2.0 * 8 * 16 * 16 * 2 * 48 ** 3 and by == 2 * 48 ** 3 * (16 + 128) * 2
    assert bench.kernel_cost("vs_inorm_relu_apply", (1, 2, 884736, 8))[1] == 2.0 * 2 * 884736 * 8 * 2
    assert set(bench.MODES) == {"joint", "seg", "vae", "joint_ttt"}
