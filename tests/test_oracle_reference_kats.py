"""The reference's own unit tests (its only golden vectors for this path), restated against the oracle:
blokus/src/pieces.rs:225-301 and blokus/src/board.rs:213-225."""
T, F = True, False


def test_piece_creation(orc):  # pieces.rs:225-241
    assert orc.piece_points(0) == 1 and orc.piece_num_variants(0) == orc.gen_variants_count([[T]])
    assert orc.piece_points(1) == 2 and orc.piece_num_variants(1) == orc.gen_variants_count([[T, T]])
    assert orc.piece_points(2) == 3 and orc.piece_num_variants(2) == 4      # Right
    assert orc.piece_points(19) == 5 and orc.piece_num_variants(19) == 8    # Crazy


def test_variant_creation(orc):  # pieces.rs:253-264
    v = orc.variant_new([[T]])
    assert v["variant"] == [True] and v["offsets"] == [0] and v["width"] == 1
    v = orc.variant_new([[T], [T]])
    assert len(v["variant"]) == 20 + 1 and v["offsets"] == [0, 20] and v["width"] == 1


def test_piece_rotation(orc):  # pieces.rs:267-275
    assert orc.rotate([[T, T]]) == [[T], [T]]
    assert orc.rotate([[T, T], [T, F]]) == [[T, T], [F, T]]


def test_piece_flip(orc):  # pieces.rs:278-286
    assert orc.flip([[T, T]]) == [[T, T]]
    assert orc.flip([[T, T], [T, F]]) == [[T, T], [F, T]]


def test_piece_variants(orc):  # pieces.rs:289-301
    assert orc.gen_variants_count([[T, T]]) == 2
    assert orc.gen_variants_count([[T, T], [T, F]]) == 4
    assert orc.gen_variants_count([[T, T, T], [T, F, F]]) == 8


def test_get_shape(orc):  # pieces.rs:244-250 — the shape survives PieceVariant::new
    assert orc.variant_new([[T, T]])["variant"] == [True, True]
    v = orc.variant_new([[T, T], [T, F]])
    assert v["variant"][:2] == [True, True] and v["variant"][20:22] == [True, False] and v["width"] == 2


def test_board_creation(orc):  # board.rs:213-217
    assert orc.fresh_board_len() == 400


def test_is_valid_move(orc):  # board.rs:219-225
    assert orc.fresh_board_is_valid(0, [[T, T]], 0) is True
    assert orc.fresh_board_is_valid(0, [[T, T]], 19) is False
