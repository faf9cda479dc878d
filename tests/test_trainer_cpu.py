"""CPU: host-side control flow of MultimodalTrainer.train_epoch (no kernels run: train_step is replaced).
The reference loop (model/trainer.py:50-166) reports a failing batch and carries on (`except Exception: continue`,
:162-164) and returns total_loss / len(dataloader) (:166)."""
import pytest
import torch


def _trainer():
    import multimodal_av_model_b200 as pkg
    from multimodal_av_model_b200.synthetic import CharTokenizer
    lin = lambda: torch.nn.Linear(2, 2)
    tr = pkg.MultimodalTrainer(lin(), lin(), lin(), lin(), CharTokenizer(30), device="cpu")
    tr.verbose = False
    return tr


def test_train_epoch_average_and_error_policy(capsys):
    tr = _trainer()
    seen = []

    def step(batch):
        seen.append(batch["x"])
        if batch["x"] < 0:
            raise RuntimeError("boom")
        return torch.tensor(float(batch["x"]))
    tr.train_step = step
    avg = tr.train_epoch([{"x": 1}, {"x": -1}, {"x": 4}, {"x": 7}])
    assert seen == [1, -1, 4, 7]                       # every batch is visited once, in order
    assert avg == pytest.approx((1 + 4 + 7) / 4)       # divided by len(dataloader), like the reference
    assert tr.last_epoch_steps == 3
    assert "Error at batch 1: boom" in capsys.readouterr().out


def test_train_epoch_empty_and_generator_errors_propagate():
    tr = _trainer()
    tr.train_step = lambda b: torch.tensor(1.0)
    assert tr.train_epoch([]) == 0.0 and tr.last_epoch_steps == 0

    class Loader:
        def __len__(self):
            return 2

        def __iter__(self):
            yield {"x": 1}
            raise OSError("worker died")               # raised by the iterator, outside the reference's try block
    with pytest.raises(OSError):
        tr.train_epoch(Loader())


def test_stage_is_idempotent_and_hot_path_has_no_cpu_fallback():
    from multimodal_av_model_b200.synthetic import make_batch
    from multimodal_av_model_b200.trainer import StagedBatch
    tr = _trainer()
    b = make_batch(pairs=1, seconds=0.2, t_v=4, l_range=(2, 3))
    s = tr.stage(b)
    assert isinstance(s, StagedBatch) and tr.stage(s) is s
    assert s["lips"][0]().shape == (1, 1, 4, 96, 96) and len(s["enc_kw"]) == 2
