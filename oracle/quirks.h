/* ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or execute it.
 *
 * Quirk flags of the CPU restatement of AnimatedRNG/alphazero-rs.
 * A set bit = the LITERAL behaviour of the reference; cleared = the corrected one.
 *   profile "reference" = AZO_PROFILE_REFERENCE (all literal)
 *   profile "sane"      = 0
 */
#ifndef AZO_QUIRKS_H
#define AZO_QUIRKS_H

/* Q1  connect_four_game.rs:114,129 — horizontal windows start only at cols 0..2,
 *     vertical windows only at rows 0..1 (exclusive ranges in the reference). */
#define AZO_Q1_WIN_RANGE_LITERAL 1u
/* Q2  async_mcts.rs:353,361-370 — the same v is applied at every level of the backup. */
#define AZO_Q2_BACKUP_NO_ALTERNATE 2u
/* Q3  node.rs:85-91 — non-negative backups add incr+1 to W (two's-complement slip). */
#define AZO_Q3_POS_BACKUP_PLUS_ONE 4u
/* Q4  coach.rs:146-153 — label +1 for the player to move at the end, -1 otherwise. */
#define AZO_Q4_VLABEL_LITERAL 8u
#define AZO_PROFILE_REFERENCE 15u
#define AZO_PROFILE_SANE 0u

/* evaluators */
#define AZO_EVAL_UNIFORM 0  /* examples/connect_four.rs:26-42 DumbConnectFourNnet */
#define AZO_EVAL_HASH 1     /* SURVEY.md App. B.6 (not in the reference) */
#define AZO_EVAL_CALLBACK 2 /* caller supplied predict(), mirrors NNet::predict (nnet.rs:40-44) */

#endif
