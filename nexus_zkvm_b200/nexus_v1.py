"""The reference's v1 main component as data: real column layout + an exact transcription of a subset of its chips (class docstring).
Column tables generated from /root/reference prover/src/column.rs (names and #[size] attributes, enum order) — layout data, not code."""
import numpy as np

from . import air as A

P = (1 << 31) - 1

MAIN_COLUMNS = [
    ("Pc", 4), ("PcNext", 4), ("PcNextAux", 4), ("OpA", 1), ("OpB", 1), ("OpC", 1), ("CarryFlag", 2), ("BorrowFlag", 2), ("ImmC", 1),
    ("InstrVal", 4), ("PrevCtr", 4), ("ValueA", 4), ("ValueAEffective", 4), ("ValueB", 4), ("ValueC", 4), ("IsAdd", 1), ("IsOr", 1), ("IsAnd", 1),
    ("IsXor", 1), ("IsSub", 1), ("IsSltu", 1), ("IsSlt", 1), ("IsBne", 1), ("IsBeq", 1), ("IsBltu", 1), ("IsBlt", 1), ("IsBgeu", 1), ("IsBge", 1),
    ("IsJal", 1), ("IsSb", 1), ("IsSh", 1), ("IsSw", 1), ("IsLb", 1), ("IsLh", 1), ("IsLbu", 1), ("IsLhu", 1), ("IsLw", 1), ("IsLui", 1),
    ("IsAuipc", 1), ("IsJalr", 1), ("IsSll", 1), ("IsSrl", 1), ("IsSra", 1), ("IsMul", 1), ("IsMulhu", 1), ("IsMulh", 1), ("IsMulhsu", 1),
    ("IsDivu", 1), ("IsDiv", 1), ("IsRemu", 1), ("IsRem", 1), ("IsEcall", 1), ("IsEbreak", 1), ("IsSysDebug", 1), ("IsSysMemoryAdvise", 1),
    ("IsSysHalt", 1), ("IsSysPrivInput", 1), ("IsSysCycleCount", 1), ("IsSysStackReset", 1), ("IsSysHeapReset", 1), ("IsCustomKeccak", 1),
    ("IsPadding", 1), ("Helper1", 4), ("Helper2", 4), ("Helper3", 4), ("Helper4", 4), ("SgnA", 1), ("SgnB", 1), ("SgnC", 1), ("Neq", 1),
    ("Neq12", 1), ("Neq34", 1), ("LtFlag", 1), ("RemAux", 1), ("Rem", 4), ("QtAux", 1), ("Qt", 4), ("ShiftBit1", 1), ("ShiftBit2", 1),
    ("ShiftBit3", 1), ("ShiftBit4", 1), ("ShiftBit5", 1), ("Exp1_3", 1), ("Exp", 1), ("RemDiff", 4), ("Neq12Aux", 1), ("Neq34Aux", 1),
    ("Neq12AuxInv", 1), ("Neq34AuxInv", 1), ("SraDegreeAux", 1), ("MulP1", 2), ("MulC1", 1), ("MulP3Prime", 2), ("MulC3Prime", 1),
    ("MulP3PrimePrime", 2), ("MulC3PrimePrime", 1), ("MulP5", 2), ("MulC5", 1), ("MulCarry0", 1), ("MulCarry1", 1), ("MulCarry2_0", 1),
    ("MulCarry2_1", 1), ("MulCarry3", 1), ("IsDivideByZero", 1), ("IsAZero", 1), ("IsOverflow", 1), ("Quotient", 4), ("HelperT", 4),
    ("Remainder", 4), ("HelperU", 4), ("RemainderBorrow", 1), ("HelperUBorrow", 1), ("ValueALow", 4), ("ValueAAbsBorrow", 2),
    ("ValueAAbsBorrowHigh", 2), ("ValueBAbsBorrow", 2), ("ValueCAbsBorrow", 2), ("ValueAAbs", 4), ("ValueAAbsHigh", 4), ("ValueBAbs", 4),
    ("ValueCAbs", 4), ("ValueAEffectiveFlag", 1), ("ValueAEffectiveFlagAux", 1), ("ValueAEffectiveFlagAuxInv", 1), ("Reg1Address", 1),
    ("Reg2Address", 1), ("Reg3Address", 1), ("Reg1ValPrev", 4), ("Reg2ValPrev", 4), ("Reg3ValPrev", 4), ("Reg1TsPrev", 4), ("Reg2TsPrev", 4),
    ("Reg3TsPrev", 4), ("ProgCtrPrev", 4), ("ProgCtrCur", 4), ("ProgCtrCarry", 2), ("FinalPrgMemoryCtr", 4), ("CReg1TsPrev", 4), ("CReg2TsPrev", 4),
    ("CReg3TsPrev", 4), ("CH1Minus", 2), ("CH2Minus", 2), ("CH3Minus", 2), ("RamBaseAddr", 4), ("Ram1ValCur", 1), ("Ram2ValCur", 1),
    ("Ram3ValCur", 1), ("Ram4ValCur", 1), ("Ram1ValPrev", 1), ("Ram2ValPrev", 1), ("Ram3ValPrev", 1), ("Ram4ValPrev", 1), ("Ram1TsPrev", 4),
    ("Ram2TsPrev", 4), ("Ram3TsPrev", 4), ("Ram4TsPrev", 4), ("Ram1TsPrevAux", 4), ("Ram2TsPrevAux", 4), ("Ram3TsPrevAux", 4), ("Ram4TsPrevAux", 4),
    ("OpC0_3", 1), ("OpC1_3", 1), ("OpC1_4", 1), ("OpC4_7", 1), ("OpC5_7", 1), ("OpC8_10", 1), ("OpC11", 1), ("OpC12", 1), ("OpC20", 1),
    ("OpA1_4", 1), ("OpB0_3", 1), ("OpB1_4", 1), ("OpC0", 1), ("OpC4", 1), ("OpA0", 1), ("OpB0", 1), ("OpB4", 1), ("OpC12_15", 1), ("OpC16_23", 1),
    ("OpC16_19", 1), ("OpC24_31", 1), ("PcCarry", 2), ("ValueA4_7", 4), ("ValueB4_7", 4), ("ValueC4_7", 4),
]
PREPROCESSED_COLUMNS = [
    ("IsFirst", 1), ("IsLast", 1), ("Clk", 4), ("Reg1TsCur", 4), ("Reg2TsCur", 4), ("Reg3TsCur", 4),
]
PROGRAM_COLUMNS = [
    ("PrgMemoryPc", 2), ("PrgMemoryWord", 2), ("PrgMemoryFlag", 1), ("PrgInitialPc", 4),
]


# ---- chip data transcribed from the reference (column lists are the reference's const arrays) --------------------------------
# virtual columns: prover/src/virtual_column.rs
TYPE_R_OPS = ["IsAdd", "IsSub", "IsSlt", "IsSltu", "IsXor", "IsOr", "IsAnd", "IsSll", "IsSrl", "IsSra", "IsMul", "IsMulhu", "IsDiv", "IsDivu",
              "IsRem", "IsRemu", "IsMulh", "IsMulhsu"]
IS_ALU = TYPE_R_OPS
IS_LOAD = ["IsLb", "IsLh", "IsLw", "IsLbu", "IsLhu"]
IS_TYPE_S = ["IsSb", "IsSh", "IsSw"]
IS_TYPE_SYS = ["IsEcall", "IsEbreak"]
IS_TYPE_U = ["IsLui", "IsAuipc"]
IS_TYPE_B = ["IsBeq", "IsBne", "IsBlt", "IsBge", "IsBltu", "IsBgeu"]
ALU_IMM_NO_SHIFT = ["IsAdd", "IsSlt", "IsSltu", "IsXor", "IsOr", "IsAnd"]
ALU_IMM_SHIFT = ["IsSll", "IsSrl", "IsSra"]
OP_B_FLAG = ["IsSb", "IsSh", "IsSw", "IsLb", "IsLh", "IsLw", "IsLbu", "IsLhu", "IsJalr", "IsAdd", "IsSub", "IsSlt", "IsSltu", "IsXor", "IsOr", "IsAnd",
             "IsSll", "IsSrl", "IsSra", "IsBeq", "IsBne", "IsBlt", "IsBge", "IsBltu", "IsBgeu", "IsMul", "IsEcall", "IsEbreak", "IsMulhu", "IsDivu",
             "IsRemu", "IsDiv", "IsRem", "IsMulh", "IsMulhsu"]
# cpu.rs:379-420: the flags whose sum (plus IsPadding, IsCustomKeccak) must be one
CPU_OP_FLAGS = ["IsAdd", "IsSub", "IsAnd", "IsOr", "IsXor", "IsSlt", "IsSltu", "IsBne", "IsBeq", "IsBltu", "IsBgeu", "IsBlt", "IsBge", "IsJal", "IsSb",
                "IsSh", "IsSw", "IsLui", "IsAuipc", "IsJalr", "IsLb", "IsLbu", "IsLh", "IsLhu", "IsLw", "IsSll", "IsSrl", "IsSra", "IsMul", "IsMulh",
                "IsMulhsu", "IsMulhu", "IsDiv", "IsDivu", "IsRem", "IsRemu", "IsEcall", "IsEbreak"]
# range_check/range_bool.rs:30-124
BOOL_SINGLE = ["ValueAEffectiveFlag", "ImmC", "IsAdd", "IsOr", "IsAnd", "IsXor", "IsSub", "IsSltu", "IsSlt", "IsBltu", "IsBlt", "IsBgeu", "IsBge", "IsBne",
               "IsBeq", "IsJal", "IsSb", "IsSh", "IsSw", "IsLb", "IsLh", "IsLbu", "IsLhu", "IsLw", "IsLui", "IsAuipc", "IsJalr", "IsSll", "IsSrl", "IsSra",
               "IsMul", "IsMulhu", "IsMulh", "IsMulhsu", "IsDivu", "IsRemu", "IsDiv", "IsRem", "IsEcall", "IsEbreak", "IsSysCycleCount", "IsSysDebug",
               "IsSysHalt", "IsSysHeapReset", "IsSysPrivInput", "IsSysStackReset", "IsPadding", "LtFlag", "RemAux", "SgnA", "SgnB", "SgnC", "ShiftBit1",
               "ShiftBit2", "ShiftBit3", "ShiftBit4", "ShiftBit5"]
BOOL_HALF_WORD = ["CarryFlag", "PcCarry", "CH1Minus", "CH2Minus", "CH3Minus", "ProgCtrCarry", "BorrowFlag", "ValueAAbsBorrow", "ValueBAbsBorrow",
                  "ValueCAbsBorrow", "ValueAAbsBorrowHigh"]
BOOL_TYPE_R = ["OpC4", "OpA0", "OpB0", "MulCarry0", "MulCarry2_0", "MulCarry2_1", "MulCarry3", "MulC1", "MulC3Prime", "MulC3PrimePrime", "MulC5",
               "IsDivideByZero", "IsOverflow", "IsAZero", "RemainderBorrow", "HelperUBorrow"]
BOOL_TYPE_I_NO_SHIFT = ["OpC11", "OpA0", "OpB0"]
BOOL_TYPE_I_SHIFT = ["OpC4", "OpA0", "OpB0"]
BOOL_TYPE_J = ["OpC11", "OpC20", "OpA0"]
BOOL_TYPE_B = ["OpC11", "OpC12", "OpA0", "OpB4"]
BOOL_TYPE_S = ["OpC0", "OpC11", "OpA0", "OpB4"]
# range_check/range256.rs:41-95
R256_WORDS = ["Pc", "PcNextAux", "InstrVal", "PrevCtr", "ValueA", "ValueB", "ValueC", "Reg1TsPrev", "Reg2TsPrev", "Reg3TsPrev", "Helper1", "ProgCtrCur",
              "ProgCtrPrev", "FinalPrgMemoryCtr", "CReg1TsPrev", "CReg2TsPrev", "CReg3TsPrev", "RamBaseAddr", "Ram1TsPrev", "Ram2TsPrev", "Ram3TsPrev",
              "Ram4TsPrev", "Ram1TsPrevAux", "Ram2TsPrevAux", "Ram3TsPrevAux", "Ram4TsPrevAux", "Rem", "Qt", "RemDiff", "HelperT", "HelperU", "Quotient",
              "Remainder", "ValueBAbs", "ValueCAbs", "ValueAAbs", "ValueAAbsHigh", "ValueALow"]
R256_BYTES = ["Ram1ValCur", "Ram2ValCur", "Ram3ValCur", "Ram4ValCur", "Ram1ValPrev", "Ram2ValPrev", "Ram3ValPrev", "Ram4ValPrev"]
R256_HALF_WORDS = ["MulP1", "MulP3Prime", "MulP3PrimePrime", "MulP5"]
R256_TYPE_U_BYTES = ["OpC16_23", "OpC24_31"]
# range_check/range16.rs:34-40, range8.rs:46-50, range32.rs:30
R16 = [("R", ["OpC0_3", "OpA1_4", "OpB1_4"]), ("U", ["OpC12_15", "OpA1_4"]), ("INoShift", ["OpC0_3", "OpC4_7", "OpA1_4", "OpB1_4"]),
       ("IShift", ["OpC0_3", "OpA1_4", "OpB1_4"]), ("J", ["OpC4_7", "OpC12_15", "OpC16_19", "OpA1_4"]), ("B", ["OpC1_4", "OpA1_4", "OpB0_3"]),
       ("S", ["OpC1_4", "OpA1_4", "OpB0_3"])]
R8 = [("INoShift", ["OpC8_10"]), ("J", ["OpC1_3", "OpC8_10"]), ("B", ["OpC5_7", "OpC8_10"]), ("S", ["OpC5_7", "OpC8_10"]), ("R", ["MulCarry1"])]
R32 = ["OpA", "OpB", "Reg1Address", "Reg2Address", "Reg3Address"]


class NexusV1Machine:
    """The reference's v1 main component (`MachineEval<BaseComponent>`, /root/reference prover/src/components/mod.rs:39-57) recorded as data —
    the REAL trace layout (all 347 main columns of prover/src/column.rs:22-604 in enum order, `Pc` and `IsPadding` with the [0, 1] mask of
    column.rs:17-19; the 18 preprocessed + 9 program columns of column.rs:617-662 requested in TraceEval::new's order, trace/eval.rs:22-50) and an
    exact transcription of these chips, in BaseComponent's order (machine.rs:49-79):
      CpuChip               chips/cpu.rs:318-566            (padding rule, ValueAEffectiveFlag, one-hot opcode sum, register wiring for types R/I/B/S/SYS,
                                                             next-row Pc, Pc increment with carries, halt)
      AddChip, SubChip      chips/instructions/i/add.rs:98-139, sub.rs:104-146
      BitOpChip             chips/instructions/i/bit_op.rs:380-431      (its 24 four-tuple lookups; the chip has no other constraint)
      LoadStoreChip         chips/instructions/i/load_store.rs:650-685,788-903   (its 8 five-tuple RAM lookups; NOT its 39 arithmetic constraint sites)
      ProgramMemCheckChip   chips/memory_check/program_mem_check.rs:152-200,255-537   (7 constraints + 4 eight-tuple lookups)
      RegisterMemCheckChip  chips/memory_check/register_mem_check.rs:120-200,325-395   (ValueAEffective + 6 nine-tuple lookups)
      TimestampChip         chips/memory_check/timestamp.rs:73-150
      Range8/16/32/128/256  chips/range_check/*.rs          (9 + 22 + 5 + 5 + 170 = 211 lookup fractions, numerators as upstream)
    — all 253 logup fractions of the reference's main component = 1012 interaction columns, 27 / 347 / 1012 committed columns —
      RangeBoolChip         chips/range_check/range_bool.rs:143-188   (112 booleanity constraints)
    with `finalize_logup()` (one secure column per fraction) and LOG_CONSTRAINT_DEGREE = 2 (components/mod.rs:12).  NOT transcribed (their
    columns are present and committed, their arithmetic constraints are not): DecodingCheck, LoadStore's arithmetic, and the other instruction chips
    (slt, branches, jumps, lui/auipc, shifts, M extension, syscalls, custom).  Of the 8 base extensions (machine.rs:82-91) the two
    multiplicity tables the witness below needs are built (Multiplicity256, Multiplicity32: extensions/multiplicity.rs).
    The witness is a PADDING-ONLY execution (every row IsPadding = 1, the state the reference pads short programs with): all opcode flags are zero, so
    every gated constraint holds and every column the transcribed chips leave free carries random in-range values (bytes, 5-bit register indices);
    timestamps follow TimestampChip's borrow arithmetic against the preprocessed Reg{1,2,3}TsCur = 3 clk + {1,2,3}.  The CPU checker proves and verifies
    it (tests/), so the kernels are measured on the reference's mask layout, constraint
    shapes (selector x linear combination, degree <= 4) and lookup structure instead of AddMachine's toy AIR."""

    def __init__(self, log_size):
        assert log_size >= 8
        self.log_size = log_size
        air = A.Air()
        # relations in draw order (C::draw_lookup_elements walks the chips: BitOp, LoadStore, ProgramMemCheck, RegisterMemCheck, Range8, 16, 32, 128, 256)
        self.rel_bitop = air.relation("BitOp", 4)            # (op, b nibble, c nibble, a nibble)               bit_op.rs:37
        self.rel_ls = air.relation("LoadStore", 5)           # (addr lo16, addr hi16, byte, ts lo16, ts hi16)   load_store.rs:66-67
        self.rel_prog = air.relation("ProgramCheck", 8)      # (pc lo16, pc hi16, word lo16, word hi16, counter bytes x4)
        self.rel_reg = air.relation("RegisterCheck", 9)
        self.rel8, self.rel16, self.rel32 = air.relation("Range8", 1), air.relation("Range16", 1), air.relation("Range32", 1)
        self.rel128, self.rel256 = air.relation("Range128", 1), air.relation("Range256", 1)
        self.relations = [self.rel_bitop, self.rel_ls, self.rel_prog, self.rel_reg, self.rel8, self.rel16, self.rel32, self.rel128, self.rel256]
        m = air.component(log_size, 2)
        pre, col = {}, {}
        for name, size in PREPROCESSED_COLUMNS + PROGRAM_COLUMNS:           # TraceEval::new: preprocessed ids, then program ids
            pre[name] = [m.get_preprocessed_column(f"{name}_{i}") for i in range(size)]
        nxt = {}
        for name, size in MAIN_COLUMNS:
            if name in ("Pc", "IsPadding"):                                 # Column::reads_next_row_mask
                pairs = [m.next_interaction_mask(A.ORIGINAL_TRACE_IDX, [0, 1]) for _ in range(size)]
                col[name], nxt[name] = [p[0] for p in pairs], [p[1] for p in pairs]
            else:
                col[name] = [m.next_trace_mask() for _ in range(size)]
        self.n_main = sum(s for _, s in MAIN_COLUMNS)

        def ssum(names):
            acc = col[names[0]][0]
            for n in names[1:]:
                acc = acc + col[n][0]
            return acc

        c = lambda n: col[n][0]
        imm_c = c("ImmC")
        is_type_r = (1 - imm_c) * ssum(TYPE_R_OPS)
        is_load, is_type_s, is_type_sys, is_type_u = ssum(IS_LOAD), ssum(IS_TYPE_S), ssum(IS_TYPE_SYS), ssum(IS_TYPE_U)
        is_type_b, is_type_j = ssum(IS_TYPE_B), c("IsJal")
        alu_imm_no_shift, alu_imm_shift = imm_c * ssum(ALU_IMM_NO_SHIFT), imm_c * ssum(ALU_IMM_SHIFT)
        is_type_i_no_shift = is_load + alu_imm_no_shift + c("IsJalr")
        is_type_i = is_load + c("IsJalr") + alu_imm_no_shift + alu_imm_shift
        op_b_flag = ssum(OP_B_FLAG)
        is_pc_incremented = ssum(IS_ALU) + is_load + is_type_s + is_type_sys * (1 - c("IsSysHalt")) + is_type_u + c("IsCustomKeccak")
        reg3_accessed = (is_type_s + is_type_b + is_type_r + is_type_i + is_type_u + is_type_j
                         + is_type_sys * (c("IsSysPrivInput") + c("IsSysHeapReset") + c("IsSysStackReset")))
        vtype = {"R": is_type_r, "U": is_type_u, "INoShift": is_type_i_no_shift, "IShift": alu_imm_shift, "J": is_type_j, "B": is_type_b, "S": is_type_s}

        # ---------------- CpuChip (cpu.rs:318-566) ----------------
        is_padding, next_is_padding = c("IsPadding"), nxt["IsPadding"][0]
        next_is_first = pre["IsLast"][0]
        m.add_constraint((1 - next_is_first) * is_padding * (1 - next_is_padding))
        m.add_constraint(c("ValueAEffectiveFlagAux") * c("ValueAEffectiveFlagAuxInv") - 1)
        op_a, op_b, op_c = c("OpA"), c("OpB"), c("OpC")
        m.add_constraint(op_a * c("ValueAEffectiveFlagAux") - c("ValueAEffectiveFlag"))
        m.add_constraint(ssum(CPU_OP_FLAGS) + is_padding + c("IsCustomKeccak") - 1)
        r1a, r2a, r3a = c("Reg1Address"), c("Reg2Address"), c("Reg3Address")
        m.add_constraint((is_type_r + is_type_i) * (op_b - r1a))
        m.add_constraint(is_type_r * (op_c - r2a))
        m.add_constraint((is_type_r + is_type_i) * (op_a - r3a))
        r1v, r2v, r3v = col["Reg1ValPrev"], col["Reg2ValPrev"], col["Reg3ValPrev"]
        va, vb, vc = col["ValueA"], col["ValueB"], col["ValueC"]
        for k in (0, 2):
            m.add_constraint(op_b_flag * (r1v[k] + r1v[k + 1] * 256 - (vb[k] + vb[k + 1] * 256)))
            m.add_constraint(is_type_r * (r2v[k] + r2v[k + 1] * 256 - (vc[k] + vc[k + 1] * 256)))
        is_type_b_s = is_type_b + is_type_s
        m.add_constraint(is_type_b_s * (op_b - r1a))
        m.add_constraint(is_type_b_s * (op_a - r3a))
        for k in (0, 2):
            m.add_constraint(is_type_b_s * (r3v[k] + r3v[k + 1] * 256 - va[k] - va[k + 1] * 256))
        is_sys_halt = c("IsSysHalt")
        m.add_constraint(is_type_sys * (op_b - r1a))
        m.add_constraint(is_type_sys * (op_c - r2a))
        m.add_constraint(is_type_sys * (op_a - r3a))
        pc, pc_next, pc_carry, pc_on_next_row = col["Pc"], col["PcNext"], col["PcCarry"], nxt["Pc"]
        for k in (0, 2):
            m.add_constraint((1 - next_is_first) * (1 - next_is_padding)
                             * (pc_next[k] + pc_next[k + 1] * 256 - (pc_on_next_row[k] + pc_on_next_row[k + 1] * 256)))
        m.add_constraint(is_pc_incremented * (pc_next[0] + pc_next[1] * 256 + pc_carry[0] * 65536 - (pc[0] + pc[1] * 256) - 4))
        m.add_constraint(is_pc_incremented * (pc_next[2] + pc_next[3] * 256 + pc_carry[1] * 65536 - (pc[2] + pc[3] * 256) - pc_carry[0]))
        for k in (0, 2):
            m.add_constraint(is_type_sys * is_sys_halt * (pc[k] + pc[k + 1] * 256 - (pc_next[k] + pc_next[k + 1] * 256)))
        # ---------------- AddChip / SubChip ----------------
        cf = col["CarryFlag"]
        is_add, is_sub = c("IsAdd"), c("IsSub")
        m.add_constraint(is_add * (va[0] + va[1] * 256 + cf[0] * 65536 - (vb[0] + vb[1] * 256 + vc[0] + vc[1] * 256)))
        m.add_constraint(is_add * (va[2] + va[3] * 256 + cf[1] * 65536 - (vb[2] + vb[3] * 256 + vc[2] + vc[3] * 256 + cf[0])))
        m.add_constraint(is_sub * (va[0] + va[1] * 256 - cf[0] * 65536 - (vb[0] + vb[1] * 256 - vc[0] - vc[1] * 256)))
        m.add_constraint(is_sub * (va[2] + va[3] * 256 - cf[1] * 65536 - (vb[2] + vb[3] * 256 - vc[2] - vc[3] * 256 - cf[0])))
        # ---------------- BitOpChip: lookup structure (bit_op.rs:380-431; 4 limbs x {And, Or, Xor} x {low, high nibble} = 24 fractions) ----------------
        va47, vb47, vc47 = col["ValueA4_7"], col["ValueB4_7"], col["ValueC4_7"]
        for k in range(4):
            for op_code, flag in ((1, c("IsAnd")), (2, c("IsOr")), (3, c("IsXor"))):
                m.add_to_relation(self.rel_bitop, flag, [m.const(op_code), vb[k] - vb47[k] * 16, vc[k] - vc47[k] * 16, va[k] - va47[k] * 16])
                m.add_to_relation(self.rel_bitop, flag, [m.const(op_code), vb47[k], vc47[k], va47[k]])
        # ---------------- LoadStoreChip: lookup structure (load_store.rs:650-685,788-903; 4 bytes x {subtract prev, add cur} = 8 fractions) ----------------
        ram1 = ssum(["IsSb", "IsSh", "IsSw", "IsLb", "IsLh", "IsLbu", "IsLhu", "IsLw"])
        ram2 = ssum(["IsSh", "IsSw", "IsLh", "IsLhu", "IsLw"])
        ram34 = ssum(["IsSw", "IsLw"])
        base, clk = col["RamBaseAddr"], pre["Clk"]
        for off, acc in ((0, ram1), (1, ram2), (2, ram34), (3, ram34)):
            addr_lo, addr_hi = base[0] + off + base[1] * 256, base[2] + base[3] * 256
            ts = col[f"Ram{off + 1}TsPrev"]
            m.add_to_relation(self.rel_ls, -acc, [addr_lo, addr_hi, c(f"Ram{off + 1}ValPrev"), ts[0] + ts[1] * 256, ts[2] + ts[3] * 256])
            m.add_to_relation(self.rel_ls, acc, [addr_lo, addr_hi, c(f"Ram{off + 1}ValCur"), clk[0] + clk[1] * 256, clk[2] + clk[3] * 256])
        # ---------------- ProgramMemCheckChip (program_mem_check.rs:152-200 and the four constrain_* helpers :255-537) ----------------
        is_first = pre["IsFirst"][0]
        for k in range(4):
            m.add_constraint(is_first * (pc[k] - pre["PrgInitialPc"][k]))
        pcur, pprev, pcarry = col["ProgCtrCur"], col["ProgCtrPrev"], col["ProgCtrCarry"]
        m.add_constraint((1 - is_padding) * (pcur[0] + pcur[1] * 256 + pcarry[0] * 65536 - (pprev[0] + pprev[1] * 256 + 1)))
        m.add_constraint((1 - is_padding) * (pcur[2] + pcur[3] * 256 + pcarry[1] * 65536 - (pprev[2] + pprev[3] * 256 + pcarry[0])))
        m.add_constraint(pcarry[1])
        pm_pc, pm_word, pm_flag = pre["PrgMemoryPc"], pre["PrgMemoryWord"], pre["PrgMemoryFlag"][0]
        zero = m.const(0)
        m.add_to_relation(self.rel_prog, pm_flag, pm_pc + pm_word + [zero] * 4)                       # add_initial_digest
        m.add_to_relation(self.rel_prog, -pm_flag, pm_pc + pm_word + col["FinalPrgMemoryCtr"])        # subtract_final_digest
        iv = col["InstrVal"]
        acc_tuple = [pc[0] + pc[1] * 256, pc[2] + pc[3] * 256, iv[0] + iv[1] * 256, iv[2] + iv[3] * 256]
        m.add_to_relation(self.rel_prog, -(1 - is_padding), acc_tuple + pprev)                         # subtract_access
        m.add_to_relation(self.rel_prog, 1 - is_padding, acc_tuple + pcur)                             # add_access
        # ---------------- RegisterMemCheckChip ----------------
        vae, vaef = col["ValueAEffective"], c("ValueAEffectiveFlag")
        for k in range(4):
            m.add_constraint(vae[k] - va[k] * vaef)
        for flag, addr, ts, val in ((op_b_flag, "Reg1Address", col["Reg1TsPrev"], r1v), (is_type_r, "Reg2Address", col["Reg2TsPrev"], r2v),
                                    (reg3_accessed, "Reg3Address", col["Reg3TsPrev"], r3v)):
            m.add_to_relation(self.rel_reg, -flag, [c(addr)] + ts + val)
        for flag, addr, ts, val in ((op_b_flag, "Reg1Address", pre["Reg1TsCur"], vb), (is_type_r, "Reg2Address", pre["Reg2TsCur"], vc),
                                    (reg3_accessed, "Reg3Address", pre["Reg3TsCur"], vae)):
            m.add_to_relation(self.rel_reg, flag, [c(addr)] + ts + val)
        # ---------------- TimestampChip ----------------
        chm = [col["CH1Minus"], col["CH2Minus"], col["CH3Minus"]]
        for k in range(3):
            m.add_constraint(chm[k][1])
        for k, (cprev, cur, prev) in enumerate(((col["CReg1TsPrev"], pre["Reg1TsCur"], col["Reg1TsPrev"]),
                                                (col["CReg2TsPrev"], pre["Reg2TsCur"], col["Reg2TsPrev"]),
                                                (col["CReg3TsPrev"], pre["Reg3TsCur"], col["Reg3TsPrev"]))):
            m.add_constraint(cprev[0] + cprev[1] * 256 + prev[0] + prev[1] * 256 + 1 - (chm[k][0] * 65536 + cur[0] + cur[1] * 256))
            m.add_constraint(cprev[2] + cprev[3] * 256 + prev[2] + prev[3] * 256 + chm[k][0] - (chm[k][1] * 65536 + cur[2] + cur[3] * 256))
        # ---------------- RangeCheckChip = (Range8, Range16, Range32, Range128, Range256, RangeBool) ----------------
        for t, cols_ in R8:
            for n in cols_:
                m.add_to_relation(self.rel8, vtype[t], [c(n)])
        m.add_to_relation(self.rel8, ssum(ALU_IMM_SHIFT), [col["Helper1"][0]])          # Helper1MsbChecked = IsSll + IsSrl + IsSra
        for t, cols_ in R16:
            for n in cols_:
                m.add_to_relation(self.rel16, vtype[t], [c(n)])
        for n in R32:
            m.add_to_relation(self.rel32, 1, [c(n)])
        num = c("IsSlt") + c("IsBge") + c("IsBlt")
        for n in ("Helper2", "Helper3"):
            m.add_to_relation(self.rel128, num, [col[n][3]])
        m.add_to_relation(self.rel128, c("IsJalr"), [c("QtAux")])
        m.add_to_relation(self.rel128, c("IsSra"), [col["Helper2"][0]])
        m.add_to_relation(self.rel128, c("IsLh") + c("IsLb"), [c("QtAux")])
        for n in R256_WORDS:
            for limb in col[n]:
                m.add_to_relation(self.rel256, 1, [limb])
        for n in R256_HALF_WORDS:
            for limb in col[n]:
                m.add_to_relation(self.rel256, 1, [limb])
        for n in R256_BYTES:
            m.add_to_relation(self.rel256, 1, [c(n)])
        for n in R256_TYPE_U_BYTES:
            m.add_to_relation(self.rel256, is_type_u, [c(n)])
        for n in BOOL_SINGLE:
            m.add_constraint(c(n) * (c(n) - 1))
        for n in BOOL_HALF_WORD:
            for limb in col[n]:
                m.add_constraint(limb * (limb - 1))
        for t, names in (("R", BOOL_TYPE_R), ("INoShift", BOOL_TYPE_I_NO_SHIFT), ("IShift", BOOL_TYPE_I_SHIFT), ("J", BOOL_TYPE_J),
                         ("B", BOOL_TYPE_B), ("S", BOOL_TYPE_S)):
            for n in names:
                m.add_constraint(vtype[t] * c(n) * (c(n) - 1))
        m.finalize_logup()                                                     # components/mod.rs:52-54
        # ---------------- extensions: Multiplicity256, Multiplicity32 (extensions/multiplicity.rs) ----------------
        t256 = air.component(8, 1)
        v256 = t256.get_preprocessed_column("Range256Values")
        t256.add_to_relation(self.rel256, -t256.next_trace_mask(), [v256])
        t256.finalize_logup()
        t32 = air.component(5, 1)
        v32 = t32.get_preprocessed_column("Range32Values")
        t32.add_to_relation(self.rel32, -t32.next_trace_mask(), [v32])
        t32.finalize_logup()
        self.air, self.main = air, m
        self.log_sizes = [log_size, 8, 5]
        self.words = air.serialize()

    # ---- columns ----
    def preprocessed_columns(self):
        """Tree 0 in commitment order: one (27, 2^log_size) uint8 block (every preprocessed / program column is a bit or a byte limb here) followed
        by the two table columns.  Built once and cached (the reference's PreprocessedTraces::new is host-side setup, machine.rs:144); a caller
        may replace `self._pre_cache[0]` by a pinned copy."""
        if getattr(self, "_pre_cache", None) is None:
            cols = self._preprocessed_list()
            block = np.stack([c for c in cols[:27]]).astype(np.uint8)
            self._pre_cache = [block] + cols[27:]
        return list(self._pre_cache)

    def _preprocessed_list(self):
        n = 1 << self.log_size
        cols = [None] * self.air.n_columns()[0]
        ids = self.air.preprocessed_ids

        def put(name, limbs):
            for i, v in enumerate(limbs):
                cols[ids[f"{name}_{i}"]] = v

        z = np.zeros(n, np.uint32)
        first, last = z.copy(), z.copy()
        first[0], last[n - 1] = 1, 1
        put("IsFirst", [first]); put("IsLast", [last])
        clk = np.arange(1, n + 1, dtype=np.uint64)                          # preprocessed.rs:75-100
        le = lambda x: [((x >> (8 * k)) & 0xFF).astype(np.uint32) for k in range(4)]
        put("Clk", le(clk)); put("Reg1TsCur", le(3 * clk + 1)); put("Reg2TsCur", le(3 * clk + 2)); put("Reg3TsCur", le(3 * clk + 3))
        for name, size in PROGRAM_COLUMNS:                                  # empty program: the program trace is all zero
            put(name, [z] * size)
        cols[ids["Range256Values"]] = np.arange(256, dtype=np.uint32)
        cols[ids["Range32Values"]] = np.arange(32, dtype=np.uint32)
        return cols

    def fill_main_trace(self, seed=0):
        """The padding-only witness (class docstring).  Returns the tree-1 columns in commitment order: 347 main, then the two multiplicity columns."""
        n = 1 << self.log_size
        rng = np.random.default_rng(seed)
        t = {name: [np.zeros(n, np.uint32) for _ in range(size)] for name, size in MAIN_COLUMNS}
        t["IsPadding"][0][:] = 1
        t["ValueAEffectiveFlagAux"][0][:] = 1
        t["ValueAEffectiveFlagAuxInv"][0][:] = 1
        rb = lambda: rng.integers(0, 256, n, dtype=np.uint64).astype(np.uint32)
        free_words = [w for w in R256_WORDS if w not in ("Reg1TsPrev", "Reg2TsPrev", "Reg3TsPrev", "CReg1TsPrev", "CReg2TsPrev", "CReg3TsPrev")]
        for w in free_words + R256_HALF_WORDS + R256_BYTES + ["ValueAEffective", "Reg1ValPrev", "Reg2ValPrev", "Reg3ValPrev", "PcNext"]:
            if w == "ValueAEffective":
                continue                                                    # = ValueA * flag = 0 (register_mem_check.rs:129-136)
            for limb in t[w]:
                limb[:] = rb()
        for limb in t["Pc"]:
            limb[0] = 0                                                    # is_first * (pc - PrgInitialPc) with an empty program
        for w in ("ValueA4_7", "ValueB4_7", "ValueC4_7"):                   # high nibbles of the (free) values; only looked up when a bit-op flag is set
            src = {"ValueA4_7": "ValueA", "ValueB4_7": "ValueB", "ValueC4_7": "ValueC"}[w]
            for k in range(4):
                t[w][k][:] = t[src][k] >> 4
        for w in ("OpB", "Reg1Address", "Reg2Address", "Reg3Address"):       # Range32-checked; OpA = ValueAEffectiveFlag / aux = 0
            t[w][0][:] = rng.integers(0, 32, n, dtype=np.uint64).astype(np.uint32)
        # timestamps: prev < cur = 3 clk + k, c = cur - 1 - prev with the chip's per-16-bit borrows (timestamp.rs:101-150)
        clk = np.arange(1, n + 1, dtype=np.uint64)
        for k in (1, 2, 3):
            cur = 3 * clk + k
            prev = (rng.integers(0, 1 << 62, n, dtype=np.uint64) % cur).astype(np.uint64)
            cc = cur - 1 - prev
            lo = (cc & 0xFFFF) + (prev & 0xFFFF) + 1
            b0 = (lo >> 16).astype(np.uint32)
            hi = (cc >> 16) + (prev >> 16) + b0
            assert np.all(lo - (b0.astype(np.uint64) << 16) == (cur & 0xFFFF)) and np.all(hi == (cur >> 16))
            for i in range(4):
                t[f"Reg{k}TsPrev"][i][:] = ((prev >> (8 * i)) & 0xFF).astype(np.uint32)
                t[f"CReg{k}TsPrev"][i][:] = ((cc >> (8 * i)) & 0xFF).astype(np.uint32)
            t[f"CH{k}Minus"][0][:] = b0
        cols = [limb for name, _ in MAIN_COLUMNS for limb in t[name]]
        h256, h32 = np.zeros(256, np.int64), np.zeros(32, np.int64)
        for w in R256_WORDS + R256_HALF_WORDS + R256_BYTES:
            for limb in t[w]:
                h256 += np.bincount(limb, minlength=256)
        for w in R32:
            h32 += np.bincount(t[w][0], minlength=32)
        cols.append((h256 % P).astype(np.uint32))
        cols.append((h32 % P).astype(np.uint32))
        assert len(cols) == self.air.n_columns()[1]
        return cols

    # ---- rank-local witness generation (one proof over N GPUs at sizes whose full trace no single host process should build) ----
    def fill_main_trace_shard(self, seed, first, count, out=None):
        """Main columns [first, first + count) of a padding-only witness, generated WITHOUT building the other columns: every word draws from its
        own generator keyed by (seed, word name), so any column range gives the same values whichever rank asks.  Same constraints as
        `fill_main_trace`, a different random stream.  Returns (columns, partial Range256 histogram, partial Range32 histogram): the two
        multiplicity columns are the sums of the partial histograms over a partition of the 347 columns (`multiplicity_columns`)."""
        import zlib
        n = 1 << self.log_size
        sizes = dict(MAIN_COLUMNS)
        index = [(name, k) for name, size in MAIN_COLUMNS for k in range(size)]
        assert 0 <= first and first + count <= len(index)
        ones = {"IsPadding", "ValueAEffectiveFlagAux", "ValueAEffectiveFlagAuxInv"}
        ts_words = {f"{pre}Reg{k}TsPrev": k for k in (1, 2, 3) for pre in ("", "C")}
        ts_words.update({f"CH{k}Minus": k for k in (1, 2, 3)})
        nibbles = {"ValueA4_7": "ValueA", "ValueB4_7": "ValueB", "ValueC4_7": "ValueC"}
        free = set(R256_WORDS + R256_HALF_WORDS + R256_BYTES + ["Reg1ValPrev", "Reg2ValPrev", "Reg3ValPrev", "PcNext"]) - {"ValueAEffective"} - set(ts_words)
        rng_of = lambda tag: np.random.default_rng([int(seed) & 0xFFFFFFFF, zlib.crc32(tag.encode())])
        cache = {}

        def timestamps(k):
            if ("ts", k) not in cache:
                clk = np.arange(1, n + 1, dtype=np.uint64)
                cur = 3 * clk + k
                prev = (rng_of(f"ts{k}").integers(0, 1 << 62, n, dtype=np.uint64) % cur).astype(np.uint64)
                cc = cur - 1 - prev
                b0 = (((cc & 0xFFFF) + (prev & 0xFFFF) + 1) >> 16).astype(np.uint8)
                cache[("ts", k)] = (prev, cc, b0)
            return cache[("ts", k)]

        def word(name):
            """all limbs of one word"""
            if name in cache:
                return cache[name]
            size = sizes[name]
            if name in ones:
                limbs = [np.ones(n, np.uint8) for _ in range(size)]
            elif name in ts_words:
                prev, cc, b0 = timestamps(ts_words[name])
                src = b0 if name.startswith("CH") else (cc if name.startswith("C") else prev)
                if name.startswith("CH"):    # the borrow of the low 16 bits; the second limb (high borrow) is 0 for these values
                    limbs = [b0] + [np.zeros(n, np.uint8) for _ in range(size - 1)]
                else:
                    limbs = [((src >> (8 * i)) & 0xFF).astype(np.uint8) for i in range(4)]
            elif name in nibbles:
                limbs = [v >> 4 for v in word(nibbles[name])]
            elif name in ("OpB", "Reg1Address", "Reg2Address", "Reg3Address"):
                limbs = [rng_of(name).integers(0, 32, n, dtype=np.uint8)]
            elif name in free:
                g = rng_of(name)
                limbs = [g.integers(0, 256, n, dtype=np.uint8) for _ in range(size)]      # every limb is a byte: kept as bytes until it is stored
                if name == "Pc":
                    for limb in limbs:
                        limb[0] = 0
            else:
                limbs = [np.zeros(n, np.uint8) for _ in range(size)]
            cache[name] = limbs
            return limbs

        cols, h256, h32 = [], np.zeros(256, np.int64), np.zeros(32, np.int64)
        r256 = set(R256_WORDS + R256_HALF_WORDS + R256_BYTES)
        for i, (name, k) in enumerate(index[first:first + count]):
            col = word(name)[k]
            if out is not None:                  # a preallocated (count, 2^log_size) block, e.g. pinned host memory
                out[i] = col
            else:
                cols.append(col.astype(np.uint32))
            if name in r256:
                h256 += np.bincount(col, minlength=256)
            if name in R32:
                h32 += np.bincount(col, minlength=32)
        return (out if out is not None else cols), h256, h32

    @staticmethod
    def multiplicity_columns(h256, h32):
        """The two extension multiplicity columns (tree 1's small columns) from the histograms summed over all main columns."""
        return [(np.asarray(h256, np.int64) % P).astype(np.uint32), (np.asarray(h32, np.int64) % P).astype(np.uint32)]

    def preprocessed_shard(self, first, count, out=None):
        """Columns [first, first + count) of the 27 big preprocessed columns (a (count, 2^log_size) uint32 array; `out` = a preallocated one) and
        the two table columns; only the requested columns are generated."""
        n = 1 << self.log_size
        ids = self.air.preprocessed_ids
        tables = [np.arange(256, dtype=np.uint32), np.arange(32, dtype=np.uint32)]
        assert ids["Range256Values"] == 27 and ids["Range32Values"] == 28
        if not count:
            return None, tables
        by_id = {}
        for name in ("IsFirst", "IsLast"):
            by_id[ids[f"{name}_0"]] = (name, 0)
        for name in ("Clk", "Reg1TsCur", "Reg2TsCur", "Reg3TsCur"):
            for k in range(4):
                by_id[ids[f"{name}_{k}"]] = (name, k)
        block = out if out is not None else np.empty((count, n), np.uint32)
        clk = None
        for i in range(count):
            name, k = by_id.get(first + i, (None, 0))
            if name is None:                       # program columns: all zero (empty program)
                block[i] = 0
            elif name == "IsFirst":
                block[i] = 0; block[i, 0] = 1
            elif name == "IsLast":
                block[i] = 0; block[i, n - 1] = 1
            else:
                if clk is None:
                    clk = np.arange(1, n + 1, dtype=np.uint64)
                v = clk if name == "Clk" else 3 * clk + int(name[3])
                block[i] = (v >> (8 * k)) & 0xFF
        return block, tables

    def column_log_sizes(self):
        return self.air.column_log_sizes()
