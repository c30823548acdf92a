"""The reference-facing entry point on the CUDA operator set: `obj_colorization_main.py --mode train` then `--mode inference`
(Foreground_Instance_Colorization/obj_colorization_main.py:17-156, main_procedure.py:62-242,495-621) in a scratch working
directory -- run directory layout, TF-format snapshots, summaries, CUDA-graph replay behind the CLI, restored weights
reproducing the picture."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TRAIN = ["--mode", "train", "--batch_size", "4", "--small_img", "1", "--synthetic_input", "1", "--summary_write_freq", "1",
         "--count_left_time_freq", "2"]


def _sketch_png(path):
    from PIL import Image, ImageDraw
    im = Image.new("L", (256, 256), 255)
    d = ImageDraw.Draw(im)
    d.rectangle([40, 120, 220, 180], outline=0, width=3)
    d.ellipse([60, 170, 100, 210], outline=0, width=3)
    d.ellipse([160, 170, 200, 210], outline=0, width=3)
    d.line([70, 120, 100, 80, 170, 80, 200, 120], fill=0, width=3)
    im.save(path)


def test_cli_train_then_inference(tmp_path, monkeypatch, capsys):
    import obj_colorization_main as M
    from sketchyscenecolorization_b200 import checkpoint, main_procedure
    from sketchyscenecolorization_b200.config import Config
    monkeypatch.chdir(tmp_path)
    M.main(TRAIN + ["--max_iter", "6", "--save_model_freq", "3"])
    out = capsys.readouterr().out
    assert "NaN" not in out and "Save model_5.ckpt" in out
    (stamp,) = os.listdir("outputs")
    run = tmp_path / "outputs" / stamp
    assert json.load(open(run / "log" / "param_0.json"))["batch_size"] == 4
    prefix = checkpoint.latest_checkpoint(str(run / "snapshot"))
    assert prefix.endswith("model_5.ckpt-5") and os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    recs = [json.loads(ln) for ln in open(run / "log" / "summaries.jsonl")]
    assert [r["step"] for r in recs] == list(range(6)) and all(np.isfinite(r["total_loss/g"]) and np.isfinite(r["total_loss/d"])
                                                                for r in recs)
    # ---- inference through the CLI
    os.makedirs("examples")
    _sketch_png("examples/car.png")
    M.main(["--mode", "inference", "--resume_from", stamp, "--small_img", "1", "--infer_name", "car.png", "--instruction",
            "the car is red with black windows"])
    import cv2
    res = run / "inference_results"
    pic, inp = cv2.imread(str(res / "car_output.png")), cv2.imread(str(res / "car_input.png"))
    assert pic.shape == (64, 64, 3) and inp.shape == (64, 64, 3) and pic.std() > 0
    # ---- restored weights reproduce the picture: two fresh processes' worth of build + restore + generate, same noise
    noise = torch.randn(1, 256, generator=torch.Generator().manual_seed(5)).cuda()
    a = main_procedure.inference("car.png", "the car is red with black windows", noise=noise)
    b = main_procedure.inference("car.png", "the car is red with black windows", noise=noise)
    assert a.shape == (1, 3, 64, 64) and np.isfinite(a).all() and np.abs(a - b).max() <= 1e-5
    saved = cv2.imread(str(res / "car_output.png"))[:, :, ::-1]
    want = (((np.transpose(b, (0, 2, 3, 1)) + 1) / 2.) * 255).astype(np.uint8)[0]
    assert np.array_equal(saved, want)
    # resume: continues at iteration 6 from the snapshot of iteration 5
    M.main(TRAIN + ["--max_iter", "8", "--save_model_freq", "2", "--resume_from", stamp])
    assert os.path.exists(run / "log" / "param_6.json")
    assert checkpoint.latest_checkpoint(str(run / "snapshot")).endswith("model_7.ckpt-7")
    assert Config.cuda_graphs == 1


def test_session_graph_replay_equals_eager_launches(tmp_path, monkeypatch):
    """main_procedure.TrainSession with and without CUDA graphs on the same seeded queues: the same loss curve (up to the
    ordering of fp32 atomics), and the graph session really replays graphs."""
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    from sketchyscenecolorization_b200.main_procedure import TrainSession
    from sketchyscenecolorization_b200.trainer import FgColorModel
    curves = []
    for graphs in (True, False):
        torch.manual_seed(11)
        torch.cuda.manual_seed(11)
        m = FgColorModel(CudaOps("cuda:0", torch.float32), "cuda:0", size=16, H=64, W=64)
        m.initialize(seed=2)
        s = TrainSession(m, batch_size=4, max_iter=100, lr_g=2e-4, lr_d=1e-4, use_cuda_graphs=graphs, small=True,
                         synthetic_input=True, data_base_dir=str(tmp_path))
        assert s.graphs is graphs
        curves.append([s.iteration() for _ in range(5)])
        if graphs:
            assert "graph" in s.tr._g["d"] and "graph" in s.tr._g["g"]
        s.close()
    for (ld_a, lg_a, nd, ng), (ld_b, lg_b, _, _) in zip(*curves):
        assert not nd and not ng
        assert abs(ld_a - ld_b) <= 5e-3 * abs(ld_b) and abs(lg_a - lg_b) <= 5e-3 * abs(lg_b), (curves[0], curves[1])
