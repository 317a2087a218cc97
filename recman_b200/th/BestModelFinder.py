"""Epoch callback that keeps the best model of a hyper-parameter sweep (recman/tf/BestModelFinder.py:9-63).

``model.fit(..., epoch_callback=finder)`` calls ``finder(model=, eval_results=, df_all=)`` after every epoch; the score is
the first metric of the last non-empty evaluation (validation when given, else training), lower is better - the
reference's convention with ``LogLoss`` first.  ``save_model=True`` writes the variables with ``DeepModel.save`` (a
``torch.save`` file instead of ``tf.train.Checkpoint``) and pickles hparams / feature dictionary / the sample frame.
"""
import logging
import os
import pickle

log = logging.getLogger(__name__)


class BestModelFinder:
    def __init__(self, save_model=False, out_dir="."):
        self._best_score = None
        self._best_eval_results = None
        self._model = None
        self.save_model = save_model
        self.out_dir = out_dir

    @property
    def best_score(self):
        return self._best_score

    @property
    def best_eval_results(self):
        return self._best_eval_results

    @property
    def best_model(self):
        return self._model

    def __call__(self, **kwargs):
        model, eval_results = kwargs.get("model"), kwargs.get("eval_results")
        assert model is not None and model.hparams is not None and model.feat_dict is not None
        assert model.variables is not None and eval_results is not None and "df_all" in kwargs
        results = [r for r in eval_results if r]  # (train, valid): valid is None without a validation set
        score = float(results[-1][0])
        if self._best_score is None or score < self._best_score:
            log.info("A better model is found! %s", results)
            self._best_score, self._best_eval_results, self._model = score, results, model
            if self.save_model:
                os.makedirs(self.out_dir, exist_ok=True)
                model.save(os.path.join(self.out_dir, "ckpt_model.pt"))
                for name, obj in (("hparams", dict(model.hparams)), ("feat_dict", model.feat_dict),
                                  ("df_all", kwargs["df_all"])):
                    with open(os.path.join(self.out_dir, name), "wb") as f:
                        pickle.dump(obj, f, protocol=pickle.HIGHEST_PROTOCOL)
