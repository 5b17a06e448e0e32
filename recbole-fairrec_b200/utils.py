"""Plugin discovery by naming convention (recbole/utils/utils.py:51-94): `get_model(name)` -> the class called `name` among
the fair recommenders of this package, `get_trainer(model_type, name)` -> the class called `<name>Trainer`, else the base
trainer.  In the reference FOCF and NFCF fall through to the base `Trainer`; here their loops live in `FOCFTrainer` (fused
step + fused full-sort evaluation) and `NFCFTrainer`, which the same `<name>Trainer` rule finds."""
import importlib

MODEL_MODULES = ("focf", "pfcn", "fairgo", "nfcf")        # the files of recbole/model/fair_recommender/, by family


def _package():
    return importlib.import_module(__name__.rsplit(".", 1)[0])


def get_model(model_name):
    """utils.py:51-73: ValueError for a name that is not a model of the package"""
    pkg = _package()
    for mod in MODEL_MODULES:
        module = importlib.import_module(f"{pkg.__name__}.{mod}")
        cls = getattr(module, str(model_name), None)
        if isinstance(cls, type) and hasattr(cls, "calculate_loss"):
            return cls
    raise ValueError("`model_name` [{}] is not the name of an existing model.".format(model_name))


def get_trainer(model_type, model_name):
    """utils.py:76-94: `<model_name>Trainer` when the package defines it, else the base trainer (`model_type` only selects
    among the reference's knowledge / traditional trainers, which the fairness models never use)"""
    pkg = _package()
    cls = getattr(pkg, f"{model_name}Trainer", None)
    return cls if isinstance(cls, type) else pkg.FOCFTrainer


def stopping_step(config, default=10):
    """`stopping_step` of overall.yaml:13 (10); an explicit 0 stays 0 (stop at the first non-improving evaluation)"""
    v = config["stopping_step"]
    return default if v is None else int(v)
