from .graphed_step import GraphedTrainStep

__all__ = ["GraphedTrainStep"]
