from .base import Estimator, Evaluator, Model, Pipeline, PipelineModel, Transformer  # noqa: F401
