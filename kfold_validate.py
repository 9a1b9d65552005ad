"""5-fold validation sweep with the reference's call shape (reference kfold_validate.py:88-113):
for every split and every model build the params, switch the model name, run validate_net."""
import os

from hyperpri_b200.src.Experiments.params_HyperPRI import ExpRedGreenBluePRI, ExpHyperspectralPRI
from hyperpri_b200.src.PLTrainer import validate_net

if __name__ == "__main__":
    rel_call_path = os.path.dirname(os.path.abspath(__file__))
    models = ['UNET', 'SpectralUNET', 'CubeNET']
    datasets = ['RGB', 'HSI', 'HSI']
    start_split, num_splits, n_seeds = 0, 5, 1
    curves = {}
    for run in range(start_split, num_splits):
        print(f" ********** Split {run + 1} **********")
        for m, dset in zip(models, datasets):
            for seed_idx in range(n_seeds):
                if dset.lower() == 'rgb':
                    exp_params = ExpRedGreenBluePRI(rel_call_path, split_no=run + 1, comet_logging=False)
                else:
                    exp_params = ExpHyperspectralPRI(rel_call_path, split_no=run + 1, comet_logging=False)
                exp_params.change_network_param(m, rel_call_path, run + 1, model_params=None)
                print(f"   Model: {exp_params.model_param_str}\n   Validation JSON: {exp_params.json_dir['val']}")
                curves[(run, m)] = validate_net(exp_params.get_val_data(), exp_params, save_segmaps=False)
