"""5-fold training driver with the reference's call shape (reference kfold_train.py:53-92):
edit the module-level globals, run `python kfold_train.py` (or under torchrun for data parallelism)."""
import os

import torch

from hyperpri_b200.src.Experiments.params_HyperPRI import ExpRedGreenBluePRI, ExpHyperspectralPRI
from hyperpri_b200.src.PLTrainer import train_net, validate_net

if __name__ == "__main__":
    rel_call_path = os.path.dirname(os.path.abspath(__file__))
    RANDOM_STATE = 1
    MODEL_SHARD = False      # reference: DeepSpeed ZeRO-2 for SpectralUNET; here: the pixel-parallel SpectralUNET option
                             # (every rank a row strip of every image); raises for the other models
    LOAD_CKPT = False
    DATA_AUG = False
    n_seeds, start_split, num_splits = 1, 0, 5
    dataset = "HSI"
    if "RANK" in os.environ and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    torch.manual_seed(RANDOM_STATE)
    for run in range(start_split, num_splits):
        print(f" ********** Split {run + 1} **********")
        for seed_idx in range(n_seeds):
            if dataset.lower() == 'rgb':
                exp_params = ExpRedGreenBluePRI(rel_call_path, split_no=run + 1, seed_num=seed_idx, augment=DATA_AUG,
                                                comet_logging=False)
            else:
                exp_params = ExpHyperspectralPRI(rel_call_path, split_no=run + 1, seed_num=seed_idx, comet_logging=False)
            pl_trainer = train_net(exp_params, checkpoint=LOAD_CKPT, model_parallel=MODEL_SHARD)
            if n_seeds > 1:
                validate_net(exp_params.get_val_data(), exp_params, save_segmaps=False)
        LOAD_CKPT = False
