"""Mirror of the train half of URAlgorithm that reaches the hot path
(/root/reference/src/main/scala/URAlgorithm.scala:130-171 params, :310-349 calcAll).
Everything else in URAlgorithm (popularity model, ES query building) is out of scope."""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Optional, Sequence

from .indexed_dataset import IndexedDataset
from .similarity_analysis import CcoContext, DownsamplableCrossOccurrenceDataset, SimilarityAnalysis


class DefaultURAlgoParams:
    """URAlgorithm.scala:53-57"""
    MaxEventsPerEventType = 500
    MaxCorrelatorsPerEventType = 50


@dataclass
class IndicatorParams:
    """URAlgorithm.scala:136-140"""
    name: str
    maxItemsPerUser: Optional[int] = None
    maxCorrelatorsPerItem: Optional[int] = None
    minLLR: Optional[float] = None


@dataclass
class URAlgorithmParams:
    """The subset of URAlgorithmParams (URAlgorithm.scala:142-171) that reaches the hot path."""
    eventNames: Optional[Sequence[str]] = None
    maxEventsPerEventType: Optional[int] = None
    maxCorrelatorsPerEventType: Optional[int] = None
    indicators: Optional[Sequence[IndicatorParams]] = None
    seed: Optional[int] = None
    recsModel: str = "all"
    # not an engine.json key of the reference: selects the literal Int/Int row sample rate recalled from Mahout 0.13.0's
    # sampleDownAndBinarize (every interaction of a user above maxItemsPerUser is dropped) instead of the
    # real division min(m, d) / d this build defaults to (INTEGRATION.md "Deviation to know about")
    rowRateIntDiv: bool = False

    @staticmethod
    def from_engine_json(algo_params: dict) -> "URAlgorithmParams":
        ind = algo_params.get("indicators")
        return URAlgorithmParams(
            eventNames=algo_params.get("eventNames"),
            maxEventsPerEventType=algo_params.get("maxEventsPerEventType"),
            maxCorrelatorsPerEventType=algo_params.get("maxCorrelatorsPerEventType"),
            indicators=None if ind is None else [IndicatorParams(i["name"], i.get("maxItemsPerUser"),
                                                                 i.get("maxCorrelatorsPerItem"), i.get("minLLR")) for i in ind],
            seed=algo_params.get("seed"), recsModel=algo_params.get("recsModel", "all"),
            rowRateIntDiv=bool(algo_params.get("rowRateIntDiv", False)))


def calc_all(actions: Sequence[tuple[str, IndexedDataset]], ap: URAlgorithmParams,
             ctx: CcoContext | None = None, flags: int = 0) -> list[tuple[str, IndexedDataset]]:
    """URAlgorithm.calcAll up to `cooccurrenceCorrelators` (URAlgorithm.scala:310-349): picks the global-
    params call or the per-indicator call, then zips the event names back on positionally (:349).
    `indicators(i)` is indexed by POSITION in `actions` exactly like the reference (:334-340)."""
    if ap.recsModel not in ("all", "collabFiltering", "backfill"):
        raise ValueError(f"Bad algorithm param recsModel=[{ap.recsModel}] in engine definition params, possibly a bad json "
                         "value. Use one of the available parameter values (all, collabFiltering, backfill).")
    if ap.recsModel == "backfill":
        return []  # calcPop only: no CCO (URAlgorithm.scala:296)
    seed = ap.seed if ap.seed is not None else int(time.time() * 1000)   # System.currentTimeMillis() (:325,345)
    if ap.rowRateIntDiv:
        flags |= 1   # CCO_FLAG_ROWRATE_INTDIV
    ids = [d for _, d in actions]
    if not ap.indicators:
        out = SimilarityAnalysis.cooccurrencesIDSs(
            ids, randomSeed=seed,
            maxInterestingItemsPerThing=ap.maxCorrelatorsPerEventType or DefaultURAlgoParams.MaxCorrelatorsPerEventType,
            maxNumInteractions=ap.maxEventsPerEventType or DefaultURAlgoParams.MaxEventsPerEventType, ctx=ctx, flags=flags)
    else:
        inds = ap.indicators
        datasets = [DownsamplableCrossOccurrenceDataset(
            iD, inds[i].maxItemsPerUser or DefaultURAlgoParams.MaxEventsPerEventType,
            inds[i].maxCorrelatorsPerItem or DefaultURAlgoParams.MaxCorrelatorsPerEventType, inds[i].minLLR)
            for i, iD in enumerate(ids)]
        out = SimilarityAnalysis.crossOccurrenceDownsampled(datasets, seed, ctx=ctx, flags=flags)
    return [(name, o) for (name, _), o in zip(actions, out)]
