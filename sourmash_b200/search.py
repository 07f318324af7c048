"""Search scoring protocol (the reference's JaccardSearch family, src/sourmash/search.py:15-169)."""
from enum import Enum


def calc_threshold_from_bp(threshold_bp, scaled, query_size):
    "threshold in bp -> (containment fraction of the query, minimum number of hashes)"
    threshold, n_threshold_hashes = 0.0, 0
    if threshold_bp:
        if threshold_bp < 0:
            raise TypeError("threshold_bp must be non-negative")
        n_threshold_hashes = float(threshold_bp) / scaled
        threshold = n_threshold_hashes / query_size
        if threshold > 1.0:
            raise ValueError("requested threshold_bp is unattainable with this query")
    return threshold, n_threshold_hashes


class SearchType(Enum):
    JACCARD = 1
    CONTAINMENT = 2
    MAX_CONTAINMENT = 3


class JaccardSearch:
    "Score function + threshold used by Index.find."

    def __init__(self, search_type, threshold=None):
        self.search_type = search_type
        self.require_scaled = search_type in (SearchType.CONTAINMENT, SearchType.MAX_CONTAINMENT)
        self.score_fn = {SearchType.JACCARD: self.score_jaccard, SearchType.CONTAINMENT: self.score_containment,
                         SearchType.MAX_CONTAINMENT: self.score_max_containment}[search_type]
        self.threshold = float(threshold or 0)

    def check_is_compatible(self, sig):
        if self.require_scaled and not sig.minhash.scaled:
            raise TypeError("this search requires a scaled signature")
        if sig.minhash.track_abundance:
            raise TypeError("this search cannot be done with an abund signature")

    def passes(self, score):
        return bool(score and score >= self.threshold)

    def collect(self, score, match_sig):
        return True

    @staticmethod
    def score_jaccard(query_size, shared_size, subject_size, total_size):
        return shared_size / total_size if total_size else 0

    @staticmethod
    def score_containment(query_size, shared_size, subject_size, total_size):
        return shared_size / query_size if query_size else 0

    @staticmethod
    def score_max_containment(query_size, shared_size, subject_size, total_size):
        d = min(query_size, subject_size)
        return shared_size / d if d else 0


class JaccardSearchBestOnly(JaccardSearch):
    "Ratchets the threshold up to the best score seen."

    def collect(self, score, match):
        self.threshold = max(self.threshold, score)
        return True


def make_jaccard_search_query(*, do_containment=False, do_max_containment=False, best_only=False, threshold=None):
    if do_containment and do_max_containment:
        raise TypeError("'do_containment' and 'do_max_containment' cannot both be True")
    cls = JaccardSearchBestOnly if best_only else JaccardSearch
    kind = SearchType.CONTAINMENT if do_containment else SearchType.MAX_CONTAINMENT if do_max_containment \
        else SearchType.JACCARD
    return cls(kind, threshold)


def make_containment_query(query_mh, threshold_bp, *, best_only=True):
    if not query_mh:
        raise ValueError("query is empty!?")
    if not query_mh.scaled:
        raise TypeError("query signature must be calculated with scaled")
    threshold, _ = calc_threshold_from_bp(threshold_bp, query_mh.scaled, len(query_mh))
    cls = JaccardSearchBestOnly if best_only else JaccardSearch
    return cls(SearchType.CONTAINMENT, threshold=threshold)
