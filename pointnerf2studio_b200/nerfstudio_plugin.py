"""`ns-train pointnerf-original` registration (reference: nerfstudio/studio_config.py:13-54,
pyproject.toml:20-21).  Import-guarded: Nerfstudio is an optional dependency that is not installed in
the build image; everything the hot path needs lives in model.py and runs without it.

Same TrainerConfig values as the reference: 4096 train/eval rays per batch, eval chunk 2304,
200k iterations, Adam 5e-4 ("fields") and 2e-3 ("neural_points"), both decayed by 0.1^(step/1e6).
"""
METHOD_NAME = "pointnerf-original"
TRAINER_VALUES = dict(max_num_iterations=200000, steps_per_save=25000, steps_per_eval_batch=1000,
                      steps_per_eval_image=2000, steps_per_eval_all_images=100000,
                      train_num_rays_per_batch=4096, eval_num_rays_per_batch=4096, eval_num_rays_per_chunk=2304,
                      lr_fields=0.0005, lr_neural_points=0.002, lr_decay_exp=0.1, lr_decay_iters=1000000)


def lr_lambda(step, lr_decay_exp=0.1, lr_decay_iters=1000000):
    """PointNerfScheduler (studio_utils.py:38-44)."""
    return pow(lr_decay_exp, step / lr_decay_iters)


try:  # pragma: no cover - exercised only where nerfstudio is installed
    from nerfstudio.engine.optimizers import AdamOptimizerConfig
    from nerfstudio.engine.schedulers import SchedulerConfig, Scheduler
    from nerfstudio.engine.trainer import TrainerConfig
    from nerfstudio.pipelines.base_pipeline import VanillaPipelineConfig
    from nerfstudio.plugins.types import MethodSpecification
    HAVE_NERFSTUDIO = True
except Exception:  # nerfstudio absent
    HAVE_NERFSTUDIO = False

if HAVE_NERFSTUDIO:  # pragma: no cover
    import dataclasses
    from typing import Type

    from torch.optim import lr_scheduler

    from .model import PointNerf, PointNerfConfig

    @dataclasses.dataclass
    class PointNerfSchedulerConfig(SchedulerConfig):
        _target: Type = dataclasses.field(default_factory=lambda: PointNerfScheduler)
        lr_decay_iters: int = 1000000
        lr_decay_exp: float = 0.1

    class PointNerfScheduler(Scheduler):
        config: PointNerfSchedulerConfig

        def get_scheduler(self, optimizer, lr_init):
            return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda s: lr_lambda(s, self.config.lr_decay_exp,
                                                                                 self.config.lr_decay_iters))

    def _make_spec():
        from nerfstudio.data.datamanagers.base_datamanager import VanillaDataManagerConfig
        sched = PointNerfSchedulerConfig(lr_decay_exp=0.1, lr_decay_iters=1000000)
        cfg = TrainerConfig(
            method_name=METHOD_NAME, experiment_name="pointnerf2studio",
            pipeline=VanillaPipelineConfig(
                datamanager=VanillaDataManagerConfig(train_num_rays_per_batch=4096, eval_num_rays_per_batch=4096),
                model=PointNerfConfig(_target=PointNerf, eval_num_rays_per_chunk=2304)),
            max_num_iterations=200000, steps_per_save=25000, steps_per_eval_batch=1000, steps_per_eval_image=2000,
            steps_per_eval_all_images=100000,
            optimizers={"fields": {"optimizer": AdamOptimizerConfig(lr=0.0005), "scheduler": sched},
                        "neural_points": {"optimizer": AdamOptimizerConfig(lr=0.002), "scheduler": sched}})
        return MethodSpecification(config=cfg, description="Point-NeRF per-ray hot path on B200 (sm_100a).")

    pointnerf_original = _make_spec()
