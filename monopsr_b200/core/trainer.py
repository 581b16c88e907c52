"""Training driver -- host-side mirror of src/monopsr/core/trainer.py:16-216 for the B200 engine.

Same control flow and console output as the reference loop: resume from the newest checkpoint of
`paths_config.checkpoint_dir` unless `overwrite_checkpoints` (trainer.py:122-167), save `<model_type>-<step:08d>`
every `checkpoint_interval` steps BEFORE the step runs (:176-184), print the total loss every `summary_interval`
steps (:196-207), run `max_iterations + 1 - start` steps.  The TF session, saver and summary writer are replaced by
`Engine.train_step` / `Engine.save_checkpoint` / `Engine.load_checkpoint`; `sample_fn()` stands in for
`model.create_feed_dict()` and returns one sample dict per call (model_spec.synthetic_sample keys, or raw
depth_map + instance_masks, see Engine.set_inputs)."""
import glob
import os
import re
import time


def latest_checkpoint(checkpoint_dir, model_type):
    """(prefix, step) of the newest '<model_type>-<step>.index' in the directory, or (None, 0)"""
    best = (None, 0)
    for path in glob.glob(os.path.join(checkpoint_dir, model_type + "-*.index")):
        m = re.search(r"-(\d+)\.index$", path)
        if m and (best[0] is None or int(m.group(1)) >= best[1]):
            best = (path[:-len(".index")], int(m.group(1)))
    return best


def prune_checkpoints(checkpoint_dir, model_type, keep):
    """tf.train.Saver(max_to_keep=...) behaviour: delete all but the `keep` newest checkpoints"""
    found = []
    for path in glob.glob(os.path.join(checkpoint_dir, model_type + "-*.index")):
        m = re.search(r"-(\d+)\.index$", path)
        if m:
            found.append((int(m.group(1)), path[:-len(".index")]))
    for _, prefix in sorted(found)[:max(0, len(found) - int(keep))]:
        for f in glob.glob(prefix + ".*"):
            os.remove(f)


def _broadcast_resume(prefix, step):
    """under torch.distributed every rank resumes from what rank 0 found (a rank that globs the directory a moment
    later could otherwise pick up the chief's fresh step-0 save)"""
    try:
        import torch.distributed as dist
    except ImportError:
        return prefix, step
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return prefix, step
    box = [(prefix, step)]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def train(engine, config, sample_fn, pretrained_checkpoint=None, log=print, chief=True):
    """trainer.train(model, config).  config: parse_yaml_config result (config_name, model_config.model_type,
    train_config.{max_iterations, summary_interval, checkpoint_interval, max_checkpoints_to_keep,
    overwrite_checkpoints, paths_config.checkpoint_dir}).  Returns the last total loss that was read back.
    chief=False (data-parallel ranks other than 0; the replicas are bit-identical): same steps, same restore, but no
    checkpoint files and no console output."""
    if not chief:
        log = lambda *a, **k: None
    tc = config.train_config
    model_type = config.model_config.model_type
    ckpt_dir = tc.paths_config.checkpoint_dir
    os.makedirs(ckpt_dir, exist_ok=True)
    prefix = os.path.join(ckpt_dir, model_type)
    log("Training", config.config_name)
    start = 0
    resume, step0 = (None, 0) if tc.overwrite_checkpoints else latest_checkpoint(ckpt_dir, model_type)
    resume, step0 = _broadcast_resume(resume, step0)      # data parallel: every rank follows rank 0's decision
    if resume is not None:
        engine.load_checkpoint(resume, resume=True)       # variables, Adam moments and EMA shadows (tf.train.Saver)
        start = step0
    elif pretrained_checkpoint is not None:
        engine.load_checkpoint(pretrained_checkpoint, kind="detection")
        log("Loading in Object Detection API pre-trained weights")
    else:
        log("Pre-trained weights are not being used.")
    engine.step_count = start
    log("Starting from step {} / {}".format(start, tc.max_iterations))
    last_time, last_loss = time.time(), None
    for step in range(start, tc.max_iterations + 1):
        if step % tc.checkpoint_interval == 0 and chief and not (resume is not None and step == start):
            # (the checkpoint a run resumed from is not written again)
            engine.save_checkpoint("{}-{:08d}".format(prefix, step), global_step=step)
            prune_checkpoints(ckpt_dir, model_type, tc.max_checkpoints_to_keep)
            log("{}: Step {} / {}: Checkpoint saved to {}-{:08d}".format(config.config_name, step, tc.max_iterations,
                                                                         prefix, step))
        engine.train_step(sample_fn())
        if step % tc.summary_interval == 0:
            now = time.time()
            last_loss = engine.losses()["total_loss"]          # the only device->host read of the loop
            log("{}: Step {}: Total Loss {:0.3f}, Time Elapsed {:0.3f} s".format(config.config_name, step, last_loss,
                                                                                 now - last_time))
            last_time = now
    return last_loss
