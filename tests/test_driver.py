"""The host driver (durf_b200/train_boxpose.py, mirror of the reference's main loop): gin parsing on CPU, a few real steps
with checkpoint + resume on the GPU."""
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GIN = os.path.join(HERE, "configs", "carla_dyn_like.gin")


def test_gin_files_parse_into_config_and_model():
    from durf_b200.obbpose_model import MipNerfModel
    from durf_b200.utils import Config, load_gin
    paths = [GIN] + [p for p in ("/root/reference/configs/carla_dyn.gin", "/root/reference/configs/waymo.gin") if os.path.exists(p)]
    for path in paths:
        cfg_kw, model_kw = load_gin(path)
        cfg = Config(**cfg_kw)
        model = MipNerfModel(**{k: v for k, v in model_kw.items() if k in MipNerfModel.__dataclass_fields__})
        assert cfg.batch_size == 512 and cfg.grad_max_val == 0.1 and cfg.eps_init == 3.0
        assert cfg.far == (40.0 if path.endswith("waymo.gin") else 200.0)
        assert model.num_samples == 128 and model.max_deg_point == 10 and model.no_pose_opt and model.contraction
        assert model.bg_topology() == (60, 256, 8, 4, 27, 128)


@pytest.mark.gpu
def test_driver_trains_checkpoints_and_resumes(tmp_path):
    from durf_b200 import train_boxpose
    common = ["--gin_file", GIN, "--train_dir", str(tmp_path), "--batch_size", "256", "--print_every", "2"]
    a = train_boxpose.main(common + ["--max_steps", "4", "--render_rows", "4"])
    assert a["step"] == 4 and all(l == l and l < 10 for l in a["losses"])
    assert os.path.exists(tmp_path / "checkpoint_4") and 0.0 <= a["render_mean_rgb"] <= 1.0
    # resumes at step 5 (state.step + 1); the first run replayed its steps from a CUDA graph, this one launches them from Python
    b = train_boxpose.main(common + ["--max_steps", "6", "--no_graph"])
    assert b["step"] == 6 and os.path.exists(tmp_path / "checkpoint_6")
    assert all(l == l and l < 10 for l in b["losses"])


@pytest.mark.gpu
def test_driver_trains_from_an_on_disk_scene(tmp_path):
    """--data_dir: batches from durf_b200.obbpose_dataset.Carla (N4) over a synthetic CARLA-layout scene, a few real steps and
    a render of the first held-out camera."""
    import sys
    sys.path.insert(0, HERE)
    import dataset_fixture as F
    from durf_b200 import train_boxpose
    scene = F.make_scene(str(tmp_path / "scene"))
    out = train_boxpose.main(["--gin_file", GIN, "--train_dir", str(tmp_path / "run"), "--data_dir", scene, "--batch_size", "256",
                              "--print_every", "1", "--max_steps", "3", "--render_rows", "4"])
    assert out["step"] == 3 and all(l == l and l < 10 for l in out["losses"]) and 0.0 <= out["render_mean_rgb"] <= 1.0
    # the loader's host rays of a camera are the rays durf_generate_rays makes on the device
    import numpy as np
    from durf_b200 import ops
    from durf_b200.obbpose_dataset import get_dataset
    from durf_b200.utils import Config
    wscene = F.make_scene(str(tmp_path / "wscene"), waymo=True)
    for root, cfg in ((scene, Config()), (wscene, Config(dataset_loader="waymo", far=40.0))):      # image centre / principal point
        test = get_dataset("test", root, cfg)
        cam = test.camera(0)
        dev_rays = ops.generate_rays(cam["c2w"], cam["width"], cam["height"], cam["focal"], cam["near"], cam["far"],
                                     principal_point=cam["principal_point"])
        for name, host, dev in zip(dev_rays._fields, test.rays, dev_rays):
            np.testing.assert_allclose(dev.cpu().numpy().reshape(host[0].shape), host[0], rtol=1e-6, atol=1e-6, err_msg=name)


def test_step_capturability_and_lazy_grad_norm():
    """Host logic next to the graphed train step: which models the driver may capture (the fp32 parity mode and a box-pose
    step whose BoxMLP is not 128 wide size their GEMMs from a host read), and stats['grad_norm'] formed on demand from the
    squared norm the step leaves on the device (train_boxpose.py:283)."""
    import torch
    from durf_b200.obbpose_model import MipNerfModel
    from durf_b200.train import Stats
    assert MipNerfModel(precision='bf16').step_is_capturable()
    assert not MipNerfModel(precision='fp32').step_is_capturable()
    assert MipNerfModel(precision='bf16', no_pose_opt=False, no_yaw_opt=False).step_is_capturable()
    assert not MipNerfModel(precision='bf16', no_pose_opt=False, box_net_width=256).step_is_capturable()
    assert MipNerfModel(precision='bf16').concurrent_objects is None       # decided per call: only while capturing
    st = Stats(grad_norm_sq=torch.tensor(6.25), loss=torch.tensor(1.0))
    assert float(st['grad_norm']) == 2.5 and 'grad_norm' not in st
    with pytest.raises(KeyError):
        st['no_such_stat']
