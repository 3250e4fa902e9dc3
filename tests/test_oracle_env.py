"""Pins the oracle's environment against every golden vector of the reference's
own unit tests (environment/src/lib.rs:196-427) and the known-answer tests
SURVEY.md section 8c derives from the code (overline, occupied cell, draw, ...)."""
import numpy as np

IP, DRAW, BW, WW = 0, 1, 2, 3


def test_place_stone_alternates(orc):  # lib.rs:201-252
    env = orc.Environment()
    assert env.turn == 0
    for i in range(12):
        assert env.place_stone(i) == IP
        assert env.board[i] == (1 if i % 2 == 0 else 2)
        assert env.turn == (i + 1) % 2


def test_game_ending_horizontal(orc):  # lib.rs:255-298
    env = orc.Environment()
    moves = [0, 9, 1, 10, 2, 11, 3, 12]
    for m in moves:
        assert env.place_stone(m) == IP
    assert env.place_stone(4) == BW


def test_game_ending_vertical(orc):  # lib.rs:301-344
    env = orc.Environment()
    for m in [0, 2, 9, 11, 18, 20, 27, 29]:
        assert env.place_stone(m) == IP
    assert env.place_stone(36) == BW


def test_game_ending_lt_rb(orc):  # lib.rs:347-358
    env = orc.Environment()
    for i in range(36):
        env.place_stone(i)
    assert env.place_stone(36 + 4) == BW


def test_game_ending_lb_rt(orc):  # lib.rs:361-372
    env = orc.Environment()
    for i in range(36):
        env.place_stone(i)
    assert env.place_stone(36) == BW


def test_encoding_0(orc):  # lib.rs:375-388
    env = orc.Environment()
    env.place_stone(0)
    exp = np.zeros(162, np.float32)
    exp[0] = 1
    assert np.array_equal(env.encode_board(0), exp)


def test_encoding_1_2(orc):  # lib.rs:391-426
    env = orc.Environment()
    for m in [0, 10, 2, 30]:
        env.place_stone(m)
    exp = np.zeros(162, np.float32)
    exp[[0, 21, 4, 61]] = 1
    assert np.array_equal(env.encode_board(0), exp)
    exp = np.zeros(162, np.float32)
    exp[[1, 20, 5, 60]] = 1
    assert np.array_equal(env.encode_board(1), exp)


# ---- known-answer tests derived from the code ----
def test_overline_is_not_a_win(orc):  # lib.rs:151-154 (== 5, not >= 5)
    env = orc.Environment()
    blacks = [0, 1, 2, 4, 5]
    whites = [18, 19, 20, 22, 23]
    for b, w in zip(blacks, whites):
        assert env.place_stone(b) == IP
        assert env.place_stone(w) == IP
    assert env.place_stone(3) == IP  # six in a row


def test_occupied_cell_is_none_without_mutation(orc):  # lib.rs:105-107
    env = orc.Environment()
    env.place_stone(40)
    t, l = env.turn, env.legal_move_count
    assert env.place_stone(40) is None
    assert (env.turn, env.legal_move_count) == (t, l)


DRAW_GRID = "121212121121212121212121212212121212121212121121212121212121212212121212121212121"


def draw_moves():
    """81 alternating moves filling the board with no five for either colour
    (pairs of rows shifted by one column; 41 black / 40 white)."""
    g = np.array([int(c) for c in DRAW_GRID], np.int8)
    blacks = [int(i) for i in np.flatnonzero(g == 1)]
    whites = [int(i) for i in np.flatnonzero(g == 2)]
    moves = []
    for k in range(40):
        moves += [blacks[k], whites[k]]
    return moves + [blacks[40]]


def test_draw_on_move_81(orc):  # lib.rs:160-161
    moves = draw_moves()
    env = orc.Environment()
    for m in moves[:-1]:
        assert env.place_stone(m) == IP
    assert env.legal_move_count == 1
    assert env.place_stone(moves[-1]) == DRAW
    assert env.legal_move_count == 0


WIN81_GRID = "122212121121212121112121212112121212121212121221212121212121212212121212121212121"


def win81_moves():
    """81 alternating moves; no five until black's LAST stone (cell 36) completes
    exactly five in column 0 on the move that also fills the board."""
    g = np.array([int(c) for c in WIN81_GRID], np.int8)
    blacks = [int(i) for i in np.flatnonzero(g == 1) if i != 36]
    whites = [int(i) for i in np.flatnonzero(g == 2)]
    moves = []
    for k in range(40):
        moves += [blacks[k], whites[k]]
    return moves + [36]


def test_win_on_move_81_beats_draw(orc):  # win test precedes draw test (lib.rs:151-161)
    moves = win81_moves()
    env = orc.Environment()
    for m in moves[:-1]:
        assert env.place_stone(m) == IP
    assert env.place_stone(moves[-1]) == BW
    assert env.legal_move_count == 0


def test_white_win_and_post_terminal_mutation(orc):
    env = orc.Environment()
    for b, w in zip([0, 1, 2, 3, 80], [9, 10, 11, 12, 13]):
        r1 = env.place_stone(b)
        assert r1 == IP
        r2 = env.place_stone(w)
    assert r2 == WW
    # no terminal guard: the board keeps accepting stones (lib.rs:104-113)
    assert env.place_stone(40) == IP
    assert env.board[40] == 1


def test_scan_reach_is_five_each_side(orc):  # lib.rs:119-144: runs up to 11 are measured
    env = orc.Environment()
    # black fills row 0 except col 4 (8 stones), white elsewhere, then col 4 -> run of 9: not a win
    blacks = [0, 1, 2, 3, 5, 6, 7, 8]
    whites = [18, 20, 22, 24, 26, 36, 38, 40]
    for b, w in zip(blacks, whites):
        assert env.place_stone(b) == IP
        assert env.place_stone(w) == IP
    assert env.place_stone(4) == IP


def test_nn_input_image(orc):  # encoder.rs:22-43 memory image
    env = orc.Environment()
    env.place_stone(0)
    env.place_stone(10)
    img = env.encode_nn_input(0)
    exp = np.zeros(243, np.float32)
    exp[0] = 1  # black at cell 0, perspective black -> offset 0
    exp[21] = 1  # white at cell 10 -> 2*10+1
    exp[162:] = 1  # black to move
    assert np.array_equal(img, exp)
    env.place_stone(5)  # white to move now
    img = env.encode_nn_input(0)
    exp = np.zeros(243, np.float32)
    exp[[1, 11]] = 1  # blacks at 0,5 seen from white: offset 1
    exp[20] = 1
    assert np.array_equal(img, exp)
    img_o = env.encode_nn_input(1)  # Opponent mode: perspective black, turn plane still white->0
    exp = np.zeros(243, np.float32)
    exp[[0, 10]] = 1
    exp[21] = 1
    assert np.array_equal(img_o, exp)
