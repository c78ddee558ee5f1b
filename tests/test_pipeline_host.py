"""Host-side semantics of SiftFeatureMatcher::Match (src/feature/matching.cc:749-839): which pairs reach the matcher,
which go straight to the verifier, what is deleted before queueing.  Pure host logic, no device needed."""
import numpy as np

from dagsfm_b200.pipeline import MatchCache, TwoViewGeometry, image_pair_to_pair_id, plan_match_jobs


def test_pair_id_is_order_free_and_follows_the_database_formula():
    # Database::ImagePairToPairId: kMaxNumImages * smaller + larger, kMaxNumImages = 2^31 - 1
    assert image_pair_to_pair_id(3, 7) == image_pair_to_pair_id(7, 3) == 2147483647 * 3 + 7
    assert image_pair_to_pair_id(0, 1) == 1


def test_self_pairs_and_duplicates_are_dropped():
    c = MatchCache()
    to_match, to_verify = plan_match_jobs([(1, 2), (2, 1), (3, 3), (1, 2), (2, 3)], c)
    assert to_match == [(1, 2), (2, 3)] and to_verify == []


def test_existing_results_follow_the_reference_rules():
    c = MatchCache()
    m = np.array([[0, 1], [2, 3]], np.uint32)
    c.WriteMatches(1, 2, m); c.WriteTwoViewGeometry(1, 2, TwoViewGeometry())          # both exist -> skipped
    c.WriteMatches(3, 4, m)                                                             # matches only -> verifier queue
    c.WriteTwoViewGeometry(5, 6, TwoViewGeometry())                                     # inliers only -> deleted, recomputed
    to_match, to_verify = plan_match_jobs([(2, 1), (4, 3), (5, 6), (7, 8)], c)
    assert to_match == [(5, 6), (7, 8)]
    assert [(a, b) for a, b, _ in to_verify] == [(4, 3)] and to_verify[0][2] is m
    assert c.ExistsMatches(1, 2) and c.ExistsInlierMatches(1, 2)                        # untouched
    assert not c.ExistsMatches(3, 4) and not c.ExistsInlierMatches(5, 6)                # deleted before queueing
    assert (c.deleted_matches, c.deleted_inlier_matches) == (1, 1)


def test_empty_pair_list():
    assert plan_match_jobs([], MatchCache()) == ([], [])
